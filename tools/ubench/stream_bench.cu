// stream_bench.cu -- B200 micro-benchmarks behind the sample-stationary encoder kernel design:
//   A. per-CTA / whole-chip TMA ingest rate (L2/HBM -> shared memory) as a function of the number of CTAs,
//      ring depth and stage size, bulk (1-D) and tensor-map (2-D box, 128-byte swizzle) copies;
//      every cluster streams the SAME per-rank region (the access pattern of 16 samples reading one weight set)
//   B. distributed-shared-memory pull bandwidth inside an 8-CTA cluster (reduce-scatter / all-gather of activations)
//   C. cluster co-residency (cudaOccupancyMaxActiveClusters) for 8-CTA clusters with ~220 KB of shared memory
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench/build/stream_bench tools/ubench/stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) __trap();
  }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------- A: stream
// mode 0: 1-D bulk copies of stage_bytes; mode 1: 2-D tensor boxes (64 x 128 bf16 = 16 KB each) over a [rows, 768] matrix
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* src,
                                                      size_t per_rank_bytes, int iters, int stages, int stage_bytes,
                                                      int mode, int csize, unsigned long long* times) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + stages * stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = blockIdx.x % csize;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 8 * (stages + s), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned long long t0 = 0;
  if (threadIdx.x == 0) t0 = gtime();
  if (warp == 0 && lane == 0) {
    const int rb_per_rank = (int)(per_rank_bytes / (12 * 16384));   // row blocks of 128 rows x 768 cols
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      mbar_wait(bars + 8 * (stages + s), ph ^ 1u);
      mbar_expect_tx(bars + 8 * s, stage_bytes);
      const uint32_t dst = base + s * stage_bytes;
      if (mode == 0) {
        const size_t off = (size_t)rank * per_rank_bytes + ((size_t)i * stage_bytes) % per_rank_bytes;
        bulk_load(dst, src + off, stage_bytes, bars + 8 * s);
      } else {
        const int nbox = stage_bytes / 16384;
        for (int b = 0; b < nbox; ++b) {
          const int t = i * nbox + b;
          const int rb = (int)rank * rb_per_rank + (t / 12) % rb_per_rank;
          tma_load_2d(dst + b * 16384, &tm, bars + 8 * s, (t % 12) * 64, rb * 128);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      mbar_wait(bars + 8 * s, ph);
      mbar_arrive(bars + 8 * (stages + s));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    times[2 * blockIdx.x] = t0;
    times[2 * blockIdx.x + 1] = gtime();
  }
}

// ---------------------------------------------------------------- B: DSMEM pull
// every CTA of an 8-cluster holds `slice_bytes * 8` of data; CTA r pulls slice r from each of the 7 peers (reduce-scatter
// traffic pattern) with 128-bit ld.shared::cluster loads and adds it up.
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int slice_bytes, int reps, float* sink, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  float4* mine = reinterpret_cast<float4*>(smem_raw);
  const uint32_t rank = cluster_rank();
  const int nvec = slice_bytes / 16;
  for (int i = threadIdx.x; i < nvec * 8; i += blockDim.x) mine[i] = make_float4(1.f, 2.f, 3.f, (float)rank);
  cluster_sync_all();
  float4 acc = make_float4(0, 0, 0, 0);
  long long c0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    for (int pr = 1; pr < 8; ++pr) {
      const uint32_t peer = (rank + pr) & 7;
      const uint32_t local = smem_u32(mine + (size_t)rank * nvec);
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        float4 v;
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(remote + i * 16));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    cluster_sync_all();
  }
  long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// same, but PUSH: every CTA writes its peers' slices into their receive buffers (st.shared::cluster)
__global__ void __launch_bounds__(256, 1) dsmem_push_kernel(int slice_bytes, int reps, float* sink, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  float4* recv = reinterpret_cast<float4*>(smem_raw);   // [8][nvec]
  const uint32_t rank = cluster_rank();
  const int nvec = slice_bytes / 16;
  for (int i = threadIdx.x; i < nvec * 8; i += blockDim.x) recv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  cluster_sync_all();
  long long c0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    for (int pr = 1; pr < 8; ++pr) {
      const uint32_t peer = (rank + pr) & 7;
      const uint32_t local = smem_u32(recv + (size_t)rank * nvec);
      uint32_t remote;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
        asm volatile("st.shared::cluster.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(remote + i * 16), "f"(1.f), "f"(2.f), "f"((float)rep), "f"((float)rank) : "memory");
      }
    }
    cluster_sync_all();
  }
  long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
  float s = 0;
  for (int i = threadIdx.x; i < nvec * 8; i += blockDim.x) s += recv[i].x + recv[i].w;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename K, typename... Args>
static cudaError_t launch_cluster(K kern, int grid, int block, size_t smem, int csize, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sms %d clock %d kHz l2 %d MB\n", prop.name, prop.multiProcessorCount, prop.clockRate, prop.l2CacheSize >> 20);
  const int csize = 8;
  const size_t per_rank = (size_t)85 * 12 * 16384;   // 16.7 MB: one CTA's weight bytes of a 12-layer forward
  const size_t total = per_rank * csize;             // 133 MB
  uint8_t* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  unsigned long long* times;
  CK(cudaMalloc(&times, 2 * 1024 * sizeof(unsigned long long)));
  // tensor map over [rows, 768] bf16
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)sym;
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {768, (cuuint64_t)(total / 1536)};
    cuuint64_t strides[1] = {1536};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  // C: co-residency of 8-CTA clusters with a big shared-memory footprint
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 8; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, stream_kernel, &cfg));
    printf("C: max active 8-CTA clusters at 220 KB smem/CTA: %d\n", ncl);
    attr[0].val.clusterDim.x = 16;
    cfg.gridDim = dim3(128);
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, stream_kernel, &cfg);
    printf("C: max active 16-CTA clusters at 220 KB smem/CTA: %d (%s)\n", ncl, cudaGetErrorString(e));
    cudaGetLastError();
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Cfg { int stages, stage_bytes; };
  const Cfg cfgs[] = {{6, 16384}, {12, 16384}, {4, 32768}, {6, 32768}, {3, 65536}};
  const int grids[] = {8, 16, 32, 64, 96, 128, 144};
  printf("A: mode grid stages stage_KB | ms | GB/s per CTA | TB/s chip | min/max CTA us\n");
  for (int mode = 0; mode < 2; ++mode) {
    for (const Cfg& c : cfgs) {
      for (int grid : grids) {
        const int iters = (int)(per_rank / c.stage_bytes);
        const size_t smem = (size_t)c.stages * c.stage_bytes + 1024 + 16 * c.stages + 64;
        float best = 1e9f;
        std::vector<unsigned long long> h(2 * grid);
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          CK(launch_cluster(stream_kernel, grid, 128, smem, csize, tm, (const uint8_t*)src, per_rank, iters, c.stages, c.stage_bytes, mode, csize, times));
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) {
            best = ms;
            CK(cudaMemcpy(h.data(), times, 2 * grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
          }
        }
        double mn = 1e30, mx = 0;
        for (int b = 0; b < grid; ++b) {
          double us = (double)(h[2 * b + 1] - h[2 * b]) * 1e-3;
          if (us < mn) mn = us;
          if (us > mx) mx = us;
        }
        const double bytes = (double)iters * c.stage_bytes;
        printf("A: %s %4d %2d %3d | %.3f | %.1f | %.2f | %.1f %.1f\n", mode ? "tensor" : "bulk  ", grid, c.stages, c.stage_bytes >> 10,
               best, bytes / (mx * 1e-6) * 1e-9, bytes * grid / (best * 1e-3) * 1e-12, mn, mx);
      }
    }
  }
  // B: DSMEM
  {
    float* sink;
    unsigned long long* cyc;
    CK(cudaMalloc(&sink, 128 * 256 * sizeof(float)));
    CK(cudaMalloc(&cyc, 128 * sizeof(unsigned long long)));
    CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    CK(cudaFuncSetAttribute(dsmem_push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const int slices[] = {2688, 5376, 10752};   // bf16 / fp32 [28 x 96] slices (and half)
    for (int sl : slices) {
      for (int grid : {8, 128}) {
        for (int push = 0; push < 2; ++push) {
          const int reps = 20;
          if (push) CK(launch_cluster(dsmem_push_kernel, grid, 256, (size_t)sl * 8, 8, sl, reps, sink, cyc));
          else CK(launch_cluster(dsmem_kernel, grid, 256, (size_t)sl * 8, 8, sl, reps, sink, cyc));
          CK(cudaDeviceSynchronize());
          std::vector<unsigned long long> h(grid);
          CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
          unsigned long long mx = 0;
          for (auto v : h) mx = v > mx ? v : mx;
          printf("B: %s slice %5d B x7 peers grid %3d: %.0f cycles per exchange (incl. cluster barrier) = %.1f B/clk/CTA\n",
                 push ? "push" : "pull", sl, grid, (double)mx / reps, 7.0 * sl / ((double)mx / reps));
        }
      }
    }
  }
  printf("done\n");
  return 0;
}
