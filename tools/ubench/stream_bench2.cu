// stream_bench2.cu -- second round of B200 micro-benchmarks for the sample-stationary encoder kernel:
//   A. TMA op-size scaling (bulk 16..128 KB, 3-D tensor boxes {64, 128, nkc}), L2-hot vs HBM-cold source,
//      L2 prefetch-ahead (cp.async.bulk.prefetch.L2), several issuing threads
//   B. latency of ONE 48 KB 3-D gather box {64 k, 32 rows, 12 chunks} from L2 (activation all-gather through global memory)
//   C. 8-CTA cluster reduce-scatter of a 28x768 fp32 partial: DSMEM pull with all loads in flight vs global-memory
//      (write, cluster barrier, read)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench/build/stream_bench2 tools/ubench/stream_bench2.cu -lcuda
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
__device__ __forceinline__ void bulk_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
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
// mode 0: bulk ops of op_bytes; mode 1: 3-D boxes {64, 128, op_bytes/16384} over the [rows, 768] matrix seen as
// {64, rows, 12}.  `issuers` lanes of warp 0 issue the ops round-robin (each owns stages s with s % issuers == lane).
// pf_ahead > 0: the producer also prefetches op i + pf_ahead into L2.
__global__ void __launch_bounds__(128, 1) stream_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* src,
                                                      size_t per_rank_bytes, int iters, int stages, int op_bytes, int mode,
                                                      int csize, int issuers, int pf_ahead, unsigned long long* times) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + stages * op_bytes;
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
  if (warp == 0 && lane < issuers) {
    const int rb_per_rank = (int)(per_rank_bytes / (12 * 16384));
    const int nkc = op_bytes / 16384;
    for (int i = lane; i < iters; i += issuers) {
      const int s = i % stages;
      const uint32_t ph = (uint32_t)(i / stages) & 1u;
      mbar_wait(bars + 8 * (stages + s), ph ^ 1u);
      mbar_expect_tx(bars + 8 * s, op_bytes);
      const uint32_t dst = base + s * op_bytes;
      if (mode == 0) {
        const size_t off = (size_t)rank * per_rank_bytes + ((size_t)i * op_bytes) % per_rank_bytes;
        if (pf_ahead > 0) {
          const size_t offp = (size_t)rank * per_rank_bytes + ((size_t)(i + pf_ahead) * op_bytes) % per_rank_bytes;
          bulk_prefetch(src + offp, op_bytes);
        }
        bulk_load(dst, src + off, op_bytes, bars + 8 * s);
      } else {
        const int per_rb = 12 / nkc;                     // ops per 128-row block
        const int rb = (int)rank * rb_per_rank + (i / per_rb) % rb_per_rank;
        tma_load_3d(dst, &tm, bars + 8 * s, 0, rb * 128, (i % per_rb) * nkc);
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

// ---------------------------------------------------------------- B: one gather box at a time (latency)
__global__ void __launch_bounds__(128, 1) gather_kernel(const __grid_constant__ CUtensorMap tm, int reps, int nrows_total,
                                                      unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 49152;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long c0 = clock64();
    for (int i = 0; i < reps; ++i) {
      mbar_expect_tx(bar, 49152);
      tma_load_3d(base, &tm, bar, 0, ((blockIdx.x * 32) + i * 32 * 64) % (nrows_total - 32), 0);
      mbar_wait(bar, i & 1);
    }
    cycles[blockIdx.x] = (unsigned long long)(clock64() - c0);
  }
}

// ---------------------------------------------------------------- C: reduce-scatter of a [28 x 768] fp32 partial per CTA
// DSMEM pull: CTA r reads rows of slice r (96 features x 28 tokens = 672 float4) from the 7 peers, all loads first
__global__ void __launch_bounds__(256, 1) rs_dsmem_kernel(int reps, float* sink, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  float4* mine = reinterpret_cast<float4*>(smem_raw);   // [8 slices][672]
  const uint32_t rank = cluster_rank();
  constexpr int NV = 672;
  for (int i = threadIdx.x; i < NV * 8; i += blockDim.x) mine[i] = make_float4(1.f, 2.f, 3.f, (float)rank);
  cluster_sync_all();
  float4 acc = make_float4(0, 0, 0, 0);
  long long c0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    uint32_t remote[7];
#pragma unroll
    for (int pr = 0; pr < 7; ++pr) {
      const uint32_t peer = (rank + 1 + pr) & 7;
      const uint32_t local = smem_u32(mine + (size_t)rank * NV);
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote[pr]) : "r"(local), "r"(peer));
    }
    float4 v[7][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < NV) {
#pragma unroll
        for (int pr = 0; pr < 7; ++pr)
          asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[pr][k].x), "=f"(v[pr][k].y), "=f"(v[pr][k].z), "=f"(v[pr][k].w) : "r"(remote[pr] + i * 16));
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < NV) {
#pragma unroll
        for (int pr = 0; pr < 7; ++pr) { acc.x += v[pr][k].x; acc.y += v[pr][k].y; acc.z += v[pr][k].z; acc.w += v[pr][k].w; }
      }
    }
    cluster_sync_all();
  }
  long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
// global-memory version: write the whole partial (8 slices x 672 float4), cluster barrier, read slice `rank` of 7 peers
__global__ void __launch_bounds__(256, 1) rs_global_kernel(int reps, float4* ws, float* sink, unsigned long long* cycles) {
  const uint32_t rank = cluster_rank();
  const int cluster = blockIdx.x / 8;
  constexpr int NV = 672;
  float4* mine = ws + ((size_t)cluster * 8 + rank) * (8 * NV);
  float4 acc = make_float4(0, 0, 0, 0);
  cluster_sync_all();
  long long c0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    for (int i = threadIdx.x; i < NV * 8; i += blockDim.x) mine[i] = make_float4(1.f, 2.f, (float)rep, (float)rank);
    __threadfence();
    cluster_sync_all();
    float4 v[7][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < NV) {
#pragma unroll
        for (int pr = 0; pr < 7; ++pr) {
          const uint32_t peer = (rank + 1 + pr) & 7;
          v[pr][k] = __ldcg(ws + ((size_t)cluster * 8 + peer) * (8 * NV) + (size_t)rank * NV + i);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < NV) {
#pragma unroll
        for (int pr = 0; pr < 7; ++pr) { acc.x += v[pr][k].x; acc.y += v[pr][k].y; acc.z += v[pr][k].z; acc.w += v[pr][k].w; }
      }
    }
    cluster_sync_all();
  }
  long long c1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(c1 - c0);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
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

static EncodeTiledFn enc;
static CUtensorMap make3d(void* base, uint64_t rows, int box_rows, int box_kc) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, rows, 12};
  cuuint64_t strides[2] = {1536, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_kc};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d (box %d x %d)\n", (int)r, box_rows, box_kc); exit(1); }
  return tm;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s sms %d\n", prop.name, prop.multiProcessorCount);
  const int csize = 8;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
  enc = (EncodeTiledFn)sym;
  const size_t cold_rank = (size_t)85 * 12 * 16384;   // 16.7 MB per rank, 133 MB total: misses L2
  const size_t hot_rank = (size_t)5 * 12 * 16384;     // 0.98 MB per rank, 7.9 MB total: L2 resident
  const size_t total = cold_rank * csize;
  uint8_t* src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  unsigned long long* times;
  CK(cudaMalloc(&times, 2 * 1024 * sizeof(unsigned long long)));
  CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  struct Cfg { int mode, stages, op_bytes, issuers, pf; };
  const Cfg cfgs[] = {
      {0, 6, 16384, 1, 0}, {0, 6, 16384, 4, 0}, {0, 6, 32768, 1, 0}, {0, 3, 65536, 1, 0}, {0, 3, 65536, 3, 0},
      {0, 1, 131072, 1, 0}, {0, 1, 196608, 1, 0}, {0, 2, 98304, 1, 0},
      {0, 6, 16384, 1, 16}, {0, 3, 65536, 1, 8}, {0, 2, 98304, 1, 6},
      {1, 6, 16384, 1, 0}, {1, 6, 32768, 1, 0}, {1, 3, 65536, 1, 0}, {1, 3, 65536, 3, 0}, {1, 2, 98304, 1, 0}, {1, 1, 196608, 1, 0}};
  const int grids[] = {8, 64, 120};
  printf("A: mode src grid stages op_KB issuers pf | ms | GB/s per CTA | TB/s chip\n");
  for (const Cfg& c : cfgs) {
    if (c.mode == 1 && (12 % (c.op_bytes / 16384)) != 0) continue;
    CUtensorMap tm = make3d(src, total / 1536, 128, c.mode == 1 ? c.op_bytes / 16384 : 1);
    for (int hot = 0; hot < 2; ++hot) {
      const size_t per_rank = hot ? hot_rank : cold_rank;
      for (int grid : grids) {
        const int iters = (int)(cold_rank / c.op_bytes);   // same bytes streamed either way
        const size_t smem = (size_t)c.stages * c.op_bytes + 1024 + 16 * c.stages + 64;
        float best = 1e9f;
        for (int rep = 0; rep < 3; ++rep) {
          CK(cudaEventRecord(e0));
          CK(launch_cluster(stream_kernel, grid, 128, smem, csize, tm, (const uint8_t*)src, per_rank, iters, c.stages, c.op_bytes,
                            c.mode, csize, c.issuers, c.pf, times));
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (ms < best) best = ms;
        }
        const double bytes = (double)iters * c.op_bytes;
        printf("A: %s %s %4d %2d %3d %d %2d | %.3f | %.1f | %.2f\n", c.mode ? "tens3d" : "bulk  ", hot ? "hot " : "cold", grid,
               c.stages, c.op_bytes >> 10, c.issuers, c.pf, best, bytes / (best * 1e-3) * 1e-9, bytes * grid / (best * 1e-3) * 1e-12);
      }
    }
  }
  // B: gather latency
  {
    CK(cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024));
    const int nrows = 448 * 16;
    CUtensorMap tg = make3d(src, nrows, 32, 12);
    unsigned long long* cyc;
    CK(cudaMalloc(&cyc, 256 * sizeof(unsigned long long)));
    for (int grid : {1, 64}) {
      for (int rep = 0; rep < 2; ++rep) {
        gather_kernel<<<grid, 128, 51 * 1024>>>(tg, 64, nrows, cyc);
        CK(cudaDeviceSynchronize());
      }
      std::vector<unsigned long long> h(grid);
      CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      unsigned long long mx = 0;
      for (auto v : h) mx = v > mx ? v : mx;
      printf("B: one 48 KB gather box {64,32,12}, grid %d: %.0f cycles per op (issue -> mbarrier complete)\n", grid, (double)mx / 64);
    }
  }
  // C: reduce-scatter
  {
    float* sink;
    unsigned long long* cyc;
    float4* ws;
    CK(cudaMalloc(&sink, 128 * 256 * sizeof(float)));
    CK(cudaMalloc(&cyc, 128 * sizeof(unsigned long long)));
    CK(cudaMalloc(&ws, (size_t)128 * 8 * 672 * sizeof(float4)));
    CK(cudaFuncSetAttribute(rs_dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int grid : {8, 64, 120}) {
      const int reps = 20;
      for (int which = 0; which < 2; ++which) {
        if (which == 0) CK(launch_cluster(rs_dsmem_kernel, grid, 256, (size_t)8 * 672 * 16, 8, reps, sink, cyc));
        else CK(launch_cluster(rs_global_kernel, grid, 256, (size_t)0, 8, reps, ws, sink, cyc));
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(grid);
        CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        unsigned long long mx = 0;
        for (auto v : h) mx = v > mx ? v : mx;
        printf("C: reduce-scatter 86 KB partial, %s, grid %3d: %.0f cycles per exchange (incl. barriers%s)\n",
               which ? "global" : "dsmem ", grid, (double)mx / reps, which ? " and the 86 KB write" : "");
      }
    }
  }
  printf("done\n");
  return 0;
}
