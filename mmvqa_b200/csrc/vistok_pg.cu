// vistok_pg.cu -- visual-token projector forward that ALSO contracts act'(.) with the feature map over the pixels
// (models/image_encoding.py:74-87, 103-113 and their autograd):
//     z[b, h, hw] = sum_c W[h, c] f[b, c, hw]
//     v[b, h]    += mean_hw act(z[b, h, hw])                               (the visual token)
//     P[b, h, c] += sum_hw act'(z[b, h, hw]) f[b, c, hw]                   (everything the weight gradient needs)
// so that   dW[h, c] = sum_b dv[b, h] / HW * P[b, h, c]   (mmvqa_vistok_dw) -- exactly the reduction the saved act' map
// [B, hidden, HW] was kept for (308 MB of bf16 written and read back per step at the 112 x 112 level, plus a K = 200 k
// split-K GEMM at the tail of the backward pass).  P does not depend on the incoming gradient, so the forward pass can
// finish the pixel contraction while the tile is still on chip:
//   * warp 0 streams [C x 64-pixel] tiles of the NCHW map through a TMA ring (128-byte swizzle).  The SAME shared-memory
//     tile is the MN-major B operand of GEMM 1 (z = W f, K = channels) and the K-major B operand of GEMM 2
//     (P += act' f^T, K = pixels): one copy, two descriptors.
//   * warp 1 issues tcgen05.mma: GEMM 1 into one of two 64-column TMEM accumulators, GEMM 2 into a third.
//   * 8 epilogue warps drain GEMM 1 (tcgen05.ld), evaluate act and act' (shared-memory SERF table), keep the running
//     row sum in registers and write act' as bf16 into shared memory in the canonical K-major SWIZZLE_128B layout --
//     the A operand of GEMM 2 -- instead of into HBM.
// Rounding is unchanged: act' is rounded to bf16 before the tensor-core contraction with fp32 accumulation, as when it
// went through HBM.  Used when only the weight gradient is needed (the feature maps are the path's input); a backbone
// that trains through the projector still takes the act'-saving path (functional.VisTokAllFn).
#include "gemm_tc_kernel.cuh"

namespace mmvqa {

constexpr int VP_BN = 64;          // pixels per tile
constexpr int VP_THREADS = 320;    // TMA warp + MMA warp + 8 epilogue warps
constexpr int VP_STAGES = 3;
constexpr int VP_REP = 8;          // SERF table replicas
constexpr int VP_D_BYTES = 16384;  // act' tile [128 x 64 pixels] bf16
constexpr int VP_TAB_BYTES = SERF_TAB_N * VP_REP * 16;
constexpr int VP_TMEM_COLS = 256;  // acc0 [0,64) | acc1 [64,128) | P [128, 128 + 64 KB)

// KB = 64-channel blocks of the map (C <= 64 KB).  KB = 1: two CTAs per SM; KB = 2 (the 80-channel level): one.
template <int KB>
struct VpCfg {
  static constexpr int A_BYTES = KB * 16384;   // W tile [128 x 64 KB] bf16
  static constexpr int B_STAGE = KB * 8192;    // f tile [64 KB channels x 64 pixels] bf16
  static constexpr int N2 = 64 * KB;           // columns of P
  static constexpr int SMEM = A_BYTES + VP_STAGES * B_STAGE + 2 * VP_D_BYTES + 1024 + 256 + VP_TAB_BYTES;
  static constexpr int CTAS_PER_SM = (2 * (SMEM + 1024) <= 233472) ? 2 : 1;
  static_assert(SMEM + 1024 <= 233472, "projector CTA exceeds the shared memory of an SM");
  static_assert(128 + N2 <= VP_TMEM_COLS, "TMEM: two z accumulators + P");
};

template <int ACT>
__device__ __forceinline__ void vp_epilogue_tile(uint32_t tmem_row, uint8_t* dbuf, int row_local, int nvalid_tile, bool row_ok,
                                                 int c_begin, float& rowsum, uint32_t tab) {
#pragma unroll 1
  for (int c = c_begin; c < c_begin + VP_BN / 2; c += 16) {
    uint32_t r[16];
    __syncwarp();
    tmem_ld16(tmem_row + (uint32_t)c, r);
    tmem_ld_wait();
    float v[16];
    const int nvalid = row_ok ? max(0, min(16, nvalid_tile - c)) : 0;
    float part = 0.0f;
    if (nvalid == 16) {     // interior chunk: no masks
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a;
        if (ACT == MMVQA_ACT_SERF) serf_both_tab<VP_REP>(tab, __uint_as_float(r[j]), a, v[j]);
        else act_both_fast<ACT>(__uint_as_float(r[j]), a, v[j]);
        part += a;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float a, d;
        const float x = j < nvalid ? __uint_as_float(r[j]) : 0.0f;
        if (ACT == MMVQA_ACT_SERF) serf_both_tab<VP_REP>(tab, x, a, d);
        else act_both_fast<ACT>(x, a, d);
        part += j < nvalid ? a : 0.0f;
        v[j] = j < nvalid ? d : 0.0f;     // pixels past the map / rows past M contribute nothing to P
      }
    }
    rowsum += part;
    // K-major SWIZZLE_128B: row r at (r / 8) * 1024 + (r % 8) * 128, its 16-byte chunk j at position j ^ (r % 8)
    uint8_t* rowp = dbuf + (row_local >> 3) * 1024 + (row_local & 7) * 128;
    const int j0 = c >> 3, sw = row_local & 7;
    *reinterpret_cast<uint4*>(rowp + (((j0) ^ sw) << 4)) =
        make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(rowp + (((j0 + 1) ^ sw) << 4)) =
        make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
  }
}

template <int KB>
__global__ void __launch_bounds__(VP_THREADS) vistok_pg_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, float* __restrict__ vis,
                                                               float* __restrict__ pgrad, int M, int N, int C, int act,
                                                               float scale, int n_tiles) {
  using Cfg = VpCfg<KB>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = a_base + Cfg::A_BYTES;
  const uint32_t d_base = b_base + VP_STAGES * Cfg::B_STAGE;
  const uint32_t bar = d_base + 2 * VP_D_BYTES;
  // barriers: a_full | b_full[S] | b_empty[S] | acc_full[2] | acc_empty[2] | d_full[2] | d_empty[2] | p_full | tmem ptr
  const uint32_t a_full = bar, b_full = bar + 8, b_empty = b_full + 8 * VP_STAGES, acc_full = b_empty + 8 * VP_STAGES,
                 acc_empty = acc_full + 16, d_full = acc_empty + 16, d_empty = d_full + 16, p_full = d_empty + 16,
                 tmem_ptr_addr = p_full + 8;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));
  uint8_t* d_gen = smem_gen + (d_base - smem_base);
  float4* tab = reinterpret_cast<float4*>(smem_gen + (bar + 256 - smem_base));
  if (act == MMVQA_ACT_SERF) serf_table_fill<VP_REP>(tab, threadIdx.x, VP_THREADS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, nsplit = gridDim.x;
  const int m0 = blockIdx.y * TC_BM, bz = blockIdx.z;
  const int my_tiles = (n_tiles - split + nsplit - 1) / nsplit;     // tiles split, split + nsplit, ...

  if (warp == 0) {
    tmem_alloc(tmem_ptr_addr, VP_TMEM_COLS);
  } else if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < VP_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, 8);      // one arrival per epilogue warp
      mbar_init(d_full + 8 * i, 8);
      mbar_init(d_empty + 8 * i, 1);
    }
    mbar_init(p_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_ptr_gen;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
      mbar_expect_tx(a_full, Cfg::A_BYTES);
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) tma_load_3d(a_base + kb * 16384, &tmA, a_full, kb * TC_BK, m0, 0);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t % VP_STAGES;
        mbar_wait(b_empty + 8 * s, ((uint32_t)(t / VP_STAGES) & 1u) ^ 1u);
        mbar_expect_tx(b_full + 8 * s, Cfg::B_STAGE);
#pragma unroll
        for (int kb = 0; kb < KB; ++kb)     // channels 64 kb .. 64 kb + 63 of the 64-pixel tile (rows past C are zero-filled)
          tma_load_3d(b_base + s * Cfg::B_STAGE + kb * 8192, &tmB, b_full + 8 * s, (split + t * nsplit) * VP_BN, kb * TC_BK, bz);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // GEMM 1: A = W K-major, B = f MN-major (pixels contiguous);  GEMM 2: A = act' K-major, B = f K-major (K = pixels)
      const uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(VP_BN >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(Cfg::N2 >> 3) << 17) |
                              ((uint32_t)(TC_BM >> 4) << 24);
      auto gemm2 = [&](int u) {
        const int bu = u & 1, su = u % VP_STAGES;
        mbar_wait(d_full + 8 * bu, ((uint32_t)u >> 1) & 1u);      // the epilogue warps have written act'(tile u)
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < VP_BN / TC_UK; ++j) {
          const uint64_t ad = make_sdesc(d_base + bu * VP_D_BYTES + j * 32, 16, 1024);
          // K-major B with N = 64 KB channel rows: the KB boxes of a stage are consecutive 8-row groups (SBO = 1024)
          const uint64_t bd = make_sdesc(b_base + su * Cfg::B_STAGE + j * 32, 16, 1024);
          umma_bf16(tmem_acc + 128u, ad, bd, idesc2, (u > 0 || j > 0) ? 1u : 0u);
        }
        umma_commit(b_empty + 8 * su);        // the pixel tile is free once BOTH contractions have read it
        umma_commit(d_empty + 8 * bu);
      };
      mbar_wait(a_full, 0);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t % VP_STAGES, buf = t & 1;
        mbar_wait(acc_empty + 8 * buf, (((uint32_t)t >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        mbar_wait(b_full + 8 * s, (uint32_t)(t / VP_STAGES) & 1u);
        tc_fence_after();
        const uint32_t sb = b_base + s * Cfg::B_STAGE;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int j = 0; j < TC_BK / TC_UK; ++j) {
            const uint64_t ad = make_sdesc(a_base + kb * 16384 + j * 32, 16, 1024);
            const uint64_t bd = make_sdesc(sb + kb * 8192 + j * 2048, 8192, 1024);
            umma_bf16(tmem_acc + (uint32_t)(buf * VP_BN), ad, bd, idesc1, (kb > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(acc_full + 8 * buf);
        if (t >= 1) gemm2(t - 1);
      }
      if (my_tiles > 0) gemm2(my_tiles - 1);
      umma_commit(p_full);
    }
  } else {
    const int g = warp & 3;
    const int half = (warp - 2) >> 2;
    const int c_begin = half * (VP_BN / 2);
    const int row_local = g * 32 + lane;
    const int m = m0 + row_local;
    const bool row_ok = m < M;
    float rowsum = 0.0f;
    const uint32_t tabh = serf_tab_handle<VP_REP>(tab);
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(acc_full + 8 * buf, ((uint32_t)t >> 1) & 1u);
      mbar_wait(d_empty + 8 * buf, (((uint32_t)t >> 1) & 1u) ^ 1u);    // GEMM 2 of tile t - 2 has read this act' buffer
      tc_fence_after();
      const uint32_t tmem_row = tmem_acc + ((uint32_t)(g * 32) << 16) + (uint32_t)(buf * VP_BN);
      const int n0 = (split + t * nsplit) * VP_BN;
      uint8_t* dbuf = d_gen + buf * VP_D_BYTES;
      const int nvt = N - n0;
      switch (act) {
        case MMVQA_ACT_SERF: vp_epilogue_tile<MMVQA_ACT_SERF>(tmem_row, dbuf, row_local, nvt, row_ok, c_begin, rowsum, tabh); break;
        case MMVQA_ACT_GELU: vp_epilogue_tile<MMVQA_ACT_GELU>(tmem_row, dbuf, row_local, nvt, row_ok, c_begin, rowsum, tabh); break;
        case MMVQA_ACT_RELU: vp_epilogue_tile<MMVQA_ACT_RELU>(tmem_row, dbuf, row_local, nvt, row_ok, c_begin, rowsum, tabh); break;
        default: vp_epilogue_tile<MMVQA_ACT_NONE>(tmem_row, dbuf, row_local, nvt, row_ok, c_begin, rowsum, tabh); break;
      }
      // accumulator back to the MMA warp; act' tile (generic-proxy stores) handed to the tensor core (async proxy)
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8 * buf) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(d_full + 8 * buf) : "memory");
      }
    }
    if (my_tiles > 0) {
      if (row_ok) atomicAdd(vis + (int64_t)bz * M + m, rowsum * scale);
      // P tile [128 rows x 64 channel columns]: this warp's lane group, its half of the columns
      mbar_wait(p_full, 0);
      tc_fence_after();
      const uint32_t prow = tmem_acc + ((uint32_t)(g * 32) << 16) + 128u;
      float* dst = pgrad + ((int64_t)bz * M + m) * C;
#pragma unroll 1
      for (int c = half * (Cfg::N2 / 2); c < (half + 1) * (Cfg::N2 / 2); c += 16) {
        if (c >= C) break;                       // warp-uniform
        uint32_t r[16];
        __syncwarp();
        tmem_ld16(prow + (uint32_t)c, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int cc = c + 4 * q;
            if (cc + 3 < C && (C & 3) == 0) {
              atomicAdd(reinterpret_cast<float4*>(dst + cc), make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                         __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (cc + e < C) atomicAdd(dst + cc + e, __uint_as_float(r[4 * q + e]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, VP_TMEM_COLS);
}

// dW[m, c] = scale * sum_b dv[b, m] * P[b, m, c]
__global__ void __launch_bounds__(256) vistok_dw_kernel(const float* __restrict__ pgrad, const float* __restrict__ dv,
                                                        float scale, float* __restrict__ dw, int B, int M, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * C) return;
  const int m = i / C;
  float acc = 0.0f;
  for (int b = 0; b < B; ++b) acc = fmaf(__ldg(dv + (int64_t)b * M + m), __ldg(pgrad + (int64_t)b * M * C + i), acc);
  dw[i] = acc * scale;
}

int tc_make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int nbatch, int64_t batch_rows,
                int box_inner, int box_rows, const char* what);

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_vistok_pgrad_supported(int M, int HW, int C) {
  return (M > 0 && C > 0 && C <= 128 && C % 8 == 0 && HW >= 2 * VP_BN) ? 1 : 0;
}

int mmvqa_vistok_fwd_pgrad(const void* W, int64_t ldw, const void* f, int64_t ldf, float* vis, float* pgrad, int M, int HW,
                           int C, int B, int act, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(W && f && vis && pgrad && B > 0, "vistok_fwd_pgrad: null pointer / empty batch");
  MMVQA_REQUIRE(mmvqa_vistok_pgrad_supported(M, HW, C), "vistok_fwd_pgrad: needs C <= 128, C %% 8 == 0 and HW >= %d (M=%d HW=%d C=%d)",
                2 * VP_BN, M, HW, C);
  MMVQA_REQUIRE(act >= MMVQA_ACT_NONE && act <= MMVQA_ACT_RELU, "vistok_fwd_pgrad: bad act %d", act);
  MMVQA_REQUIRE(ldw >= C && ldf >= HW, "vistok_fwd_pgrad: bad leading dimensions");
  MMVQA_REQUIRE((reinterpret_cast<uintptr_t>(pgrad) & 15) == 0, "vistok_fwd_pgrad: pgrad must be 16-byte aligned");
  int sm = mmvqa_device_sm();
  if (sm < 0) return sm;
  if (sm / 10 != 10) return set_err(MMVQA_ERR_ARCH, "vistok_fwd_pgrad: tcgen05 path needs sm_100, device is sm_%d", sm);
  CUtensorMap tmA, tmB;
  int rc = tc_make_map(&tmA, W, C, M, ldw, 1, 0, 64, TC_BM, "W");
  if (rc) return rc;
  rc = tc_make_map(&tmB, f, HW, C, ldf, B, C, 64, 64, "f");
  if (rc) return rc;
  const int n_tiles = (HW + VP_BN - 1) / VP_BN;
  const int mt = (M + TC_BM - 1) / TC_BM;
  const int kbn = C <= 64 ? 1 : 2;
  // ONE wave (two CTAs per SM for C <= 64, one above), at least 2 pixel tiles per CTA
  int nsplit = ((kbn == 1 ? VpCfg<1>::CTAS_PER_SM : VpCfg<2>::CTAS_PER_SM) * num_sms()) / (mt * B);
  if (nsplit > (n_tiles + 1) / 2) nsplit = (n_tiles + 1) / 2;
  if (nsplit < 1) nsplit = 1;
  dim3 grid(nsplit, mt, B);
  MMVQA_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "vistok_fwd_pgrad: grid too large");
  if (kbn == 1) {
    static bool attr_set = false;
    if (!attr_set) {
      MMVQA_CUDA(cudaFuncSetAttribute(vistok_pg_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, VpCfg<1>::SMEM));
      attr_set = true;
    }
    MMVQA_CUDA(launch_pdl(vistok_pg_kernel<1>, grid, dim3(VP_THREADS), (size_t)VpCfg<1>::SMEM, as_stream(stream), tmA, tmB, vis, pgrad,
                          M, HW, C, act, 1.0f / (float)HW, n_tiles));
  } else {
    static bool attr_set = false;
    if (!attr_set) {
      MMVQA_CUDA(cudaFuncSetAttribute(vistok_pg_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, VpCfg<2>::SMEM));
      attr_set = true;
    }
    MMVQA_CUDA(launch_pdl(vistok_pg_kernel<2>, grid, dim3(VP_THREADS), (size_t)VpCfg<2>::SMEM, as_stream(stream), tmA, tmB, vis, pgrad,
                          M, HW, C, act, 1.0f / (float)HW, n_tiles));
  }
  MMVQA_LAUNCHED("vistok_pg_kernel");
  return MMVQA_OK;
}

int mmvqa_vistok_dw(const float* pgrad, const float* dv, float scale, float* dW, int B, int M, int C, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(pgrad && dv && dW && B > 0 && M > 0 && C > 0, "vistok_dw: bad args");
  vistok_dw_kernel<<<(M * C + 255) / 256, 256, 0, as_stream(stream)>>>(pgrad, dv, scale, dW, B, M, C);
  MMVQA_LAUNCHED("vistok_dw");
  return MMVQA_OK;
}

}  // extern "C"
