// vistok.cu -- the visual-token projector kernel (models/image_encoding.py:74-87, 103-113):
//     v[b, h] = mean_hw act( sum_c W[h, c] f[b, c, hw] )                      (forward,  MODE 0)
//     G[b, h, hw] = act'( sum_c W[h, c] f[b, c, hw] ) * dv[b, h] / HW         (backward, MODE 1: recompute)
// for the pyramid levels with few channels (C <= 128), where the 1x1 conv is a K <= 128 GEMM over up to 12544
// pixels and the per-pixel activation -- not the MMA -- is the work.  Persistent, warp-specialised CTA per
// (sample, 128 output channels, pixel split):
//   * the W tile [128 x C] is fetched ONCE by TMA and stays in shared memory;
//   * warp 0 streams [C x 128-pixel] tiles of the NCHW feature map (MN-major B operand, read in place) through
//     a TMA ring; warp 1 issues tcgen05.mma into one of TWO TMEM accumulators;
//   * 8 epilogue warps drain the other accumulator (tcgen05.ld), evaluate the activation with the 2-MUFU SERF
//     and either keep a running row sum in registers (forward: one atomic per row per CTA, the [B,768,H,W] map
//     is never written) or store G as bf16 (backward);
//   so the MMA / TMA of pixel tile t+1 overlap the activation math of tile t and no per-tile CTA start-up is paid.
// Reached through mmvqa_gemm (EPI_ACT_ROWSUM / EPI_DACT_SCALE with K <= 128, K-major A, MN-major batched B).
#include "gemm_tc_kernel.cuh"

namespace mmvqa {

constexpr int VT_BN = 128;
constexpr int VT_THREADS = 320;

template <int KB, int STAGES>
struct VtCfg {
  static constexpr int A_BYTES = KB * 16384;
  static constexpr int B_STAGE = KB * 16384;
  static constexpr int REP = KB == 1 ? 8 : 4;                        // SERF table replicas (see common.cuh)
  static constexpr int TAB_BYTES = SERF_TAB_N * REP * 16;
  static constexpr int RAW = A_BYTES + STAGES * B_STAGE + 1024 + 256 + TAB_BYTES;
  static constexpr int SMEM = RAW < 80 * 1024 ? 80 * 1024 : RAW;   // never more than 2 CTAs per SM (2 x 256 TMEM columns)
  static_assert(SMEM <= 116224, "two projector CTAs must fit on one SM (227 KB)");
};

template <int MODE, int ACT, int REP>
__device__ __forceinline__ void vt_epilogue_tile(const EpiParams& p, uint32_t tmem_row, int m, int n0, int bz, bool row_ok,
                                                 int c_begin, float rscale, float& rowsum, uint32_t tab) {
#pragma unroll 1
  for (int c = c_begin; c < c_begin + VT_BN / 2; c += 16) {
    uint32_t r[16];
    __syncwarp();
    tmem_ld16(tmem_row + (uint32_t)c, r);
    tmem_ld_wait();
    const int nb = n0 + c;
    if (row_ok && nb < p.N) {
      const int nvalid = min(16, p.N - nb);
      float v[16];
      if (MODE == 0) {
        float part = 0.0f;
        if (nvalid < 16) {   // ragged last chunk of a row of pixels: columns past N hold act(0) = 0 for SERF / ReLU / GELU, but mask anyway
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j >= nvalid) r[j] = 0u;
        }
        if (p.aux_out) {   // forward that keeps act'(.) (bf16) for the backward pass
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a;
            if (ACT == MMVQA_ACT_SERF) serf_both_tab<REP>(tab, __uint_as_float(r[j]), a, v[j]);
            else act_both_fast<ACT>(__uint_as_float(r[j]), a, v[j]);
            part += a;
          }
          store16_bf16(reinterpret_cast<__nv_bfloat16*>(p.aux_out) + ((int64_t)bz * p.M + m) * p.ld_aux_out + nb, nvalid, v);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            part += ACT == MMVQA_ACT_SERF ? serf_tab<REP>(tab, __uint_as_float(r[j])) : act_fast<ACT>(__uint_as_float(r[j]));
        }
        rowsum += part;
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = (ACT == MMVQA_ACT_SERF ? dserf_tab<REP>(tab, __uint_as_float(r[j])) : dact_fast<ACT>(__uint_as_float(r[j]))) * rscale;
        store16_bf16(reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)bz * p.c_batch_stride + (int64_t)m * p.ldc + nb, nvalid, v);
      }
    }
  }
}

template <int MODE, int KB, int STAGES>
__global__ void __launch_bounds__(VT_THREADS) vistok_kernel(const __grid_constant__ CUtensorMap tmA,
                                                            const __grid_constant__ CUtensorMap tmB, EpiParams p,
                                                            int n_tiles) {
  using Cfg = VtCfg<KB, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = smem_base;
  const uint32_t b_base = smem_base + Cfg::A_BYTES;
  const uint32_t bar = b_base + STAGES * Cfg::B_STAGE;
  // barriers: a_full | b_full[S] | b_empty[S] | acc_full[2] | acc_empty[2] | tmem ptr
  const uint32_t a_full = bar, b_full = bar + 8, b_empty = b_full + 8 * STAGES, acc_full = b_empty + 8 * STAGES,
                 acc_empty = acc_full + 16, tmem_ptr_addr = acc_empty + 16;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_ptr_addr - smem_base));
  float4* tab = reinterpret_cast<float4*>(smem_gen + (bar + 256 - smem_base));     // SERF interpolation table (32 KB)
  if (p.act == MMVQA_ACT_SERF) serf_table_fill<Cfg::REP>(tab, threadIdx.x, VT_THREADS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, nsplit = gridDim.x;
  const int m0 = blockIdx.y * TC_BM, bz = blockIdx.z;
  const int my_tiles = (n_tiles - split + nsplit - 1) / nsplit;     // tiles split, split + nsplit, ...

  if (warp == 0) {
    tmem_alloc(tmem_ptr_addr, 2 * VT_BN);
  } else if (warp == 1 && lane == 0) {
    mbar_init(a_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, 8);      // one arrival per epilogue warp
    }
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
        const int s = t % STAGES;
        const uint32_t ph = (uint32_t)(t / STAGES) & 1u;
        mbar_wait(b_empty + 8 * s, ph ^ 1u);
        mbar_expect_tx(b_full + 8 * s, Cfg::B_STAGE);
        const int n0 = (split + t * nsplit) * VT_BN;
        const uint32_t sb = b_base + s * Cfg::B_STAGE;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          tma_load_3d(sb + kb * 16384, &tmB, b_full + 8 * s, n0, kb * TC_BK, bz);
          tma_load_3d(sb + kb * 16384 + 8192, &tmB, b_full + 8 * s, n0 + 64, kb * TC_BK, bz);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (0u << 15) | (1u << 16) | ((uint32_t)(VT_BN >> 3) << 17) |
                             ((uint32_t)(TC_BM >> 4) << 24);
      mbar_wait(a_full, 0);
      for (int t = 0; t < my_tiles; ++t) {
        const int s = t % STAGES, buf = t & 1;
        mbar_wait(acc_empty + 8 * buf, (((uint32_t)t >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
        mbar_wait(b_full + 8 * s, (uint32_t)(t / STAGES) & 1u);
        tc_fence_after();
        const uint32_t sb = b_base + s * Cfg::B_STAGE;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
          for (int j = 0; j < TC_BK / TC_UK; ++j) {
            const uint64_t ad = make_sdesc(a_base + kb * 16384 + j * 32, 16, 1024);
            const uint64_t bd = make_sdesc(sb + kb * 16384 + j * 2048, 8192, 1024);
            umma_bf16(tmem_acc + (uint32_t)(buf * VT_BN), ad, bd, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(b_empty + 8 * s);
        umma_commit(acc_full + 8 * buf);
      }
    }
  } else {
    const int g = warp & 3;
    const int c_begin = ((warp - 2) >> 2) * (VT_BN / 2);
    const int m = m0 + g * 32 + lane;
    const bool row_ok = m < p.M;
    float rscale = 0.0f, rowsum = 0.0f;
    const uint32_t tabh = serf_tab_handle<Cfg::REP>(tab);
    if (MODE == 1 && row_ok) rscale = __ldg(p.rowscale + (int64_t)bz * p.M + m) * p.scale;
    for (int t = 0; t < my_tiles; ++t) {
      const int buf = t & 1;
      mbar_wait(acc_full + 8 * buf, ((uint32_t)t >> 1) & 1u);
      tc_fence_after();
      const uint32_t tmem_row = tmem_acc + ((uint32_t)(g * 32) << 16) + (uint32_t)(buf * VT_BN);
      const int n0 = (split + t * nsplit) * VT_BN;
      switch (p.act) {
        case MMVQA_ACT_SERF: vt_epilogue_tile<MODE, MMVQA_ACT_SERF, Cfg::REP>(p, tmem_row, m, n0, bz, row_ok, c_begin, rscale, rowsum, tabh); break;
        case MMVQA_ACT_GELU: vt_epilogue_tile<MODE, MMVQA_ACT_GELU, Cfg::REP>(p, tmem_row, m, n0, bz, row_ok, c_begin, rscale, rowsum, tabh); break;
        case MMVQA_ACT_RELU: vt_epilogue_tile<MODE, MMVQA_ACT_RELU, Cfg::REP>(p, tmem_row, m, n0, bz, row_ok, c_begin, rscale, rowsum, tabh); break;
        default: vt_epilogue_tile<MODE, MMVQA_ACT_NONE, Cfg::REP>(p, tmem_row, m, n0, bz, row_ok, c_begin, rscale, rowsum, tabh); break;
      }
      // hand the accumulator back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8 * buf) : "memory");
    }
    if (MODE == 0 && row_ok && my_tiles > 0) atomicAdd(p.rowsum_out + (int64_t)bz * p.M + m, rowsum * p.scale);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_acc, 2 * VT_BN);
}

template <int MODE, int KB, int STAGES>
static int launch_vt(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  using Cfg = VtCfg<KB, STAGES>;
  CUtensorMap tmA, tmB;
  int rc = tc_make_map(&tmA, a->A, a->K, a->M, a->lda, 1, 0, 64, TC_BM, "A");
  if (rc) return rc;
  rc = tc_make_map(&tmB, a->B, a->N, a->K, a->ldb, a->batch, a->b_batch_rows, 64, 64, "B");
  if (rc) return rc;
  auto kern = vistok_kernel<MODE, KB, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    MMVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    attr_set = true;
  }
  const int n_tiles = (a->N + VT_BN - 1) / VT_BN;
  const int mt = (a->M + TC_BM - 1) / TC_BM;
  // two CTAs per SM (TMEM: 2 x 256 columns): ONE wave of at most 2 * SMs CTAs (rounding the split up gave 1.3 waves
  // and a 36 % idle tail at the 112 x 112 level), at least 2 pixel tiles per CTA
  int nsplit = (2 * num_sms()) / (mt * a->batch);
  if (nsplit > (n_tiles + 1) / 2) nsplit = (n_tiles + 1) / 2;
  if (nsplit < 1) nsplit = 1;
  dim3 grid(nsplit, mt, a->batch);
  MMVQA_REQUIRE(grid.z <= 65535 && grid.y <= 65535, "vistok: grid too large");
  MMVQA_CUDA(launch_pdl(kern, grid, dim3(VT_THREADS), (size_t)Cfg::SMEM, st, tmA, tmB, ep, n_tiles));
  MMVQA_LAUNCHED("vistok_kernel");
  return MMVQA_OK;
}

// true if this problem is a projector level the persistent kernel handles
bool vistok_applicable(const mmvqa_gemm_args* a) {
  return (a->epilogue == MMVQA_EPI_ACT_ROWSUM || a->epilogue == MMVQA_EPI_DACT_SCALE) && a->K <= 128 && !a->a_trans &&
         a->b_trans && a->split_k == 1 && a->a_batch_rows == 0 && a->bias == nullptr && !a->accumulate &&
         (a->epilogue == MMVQA_EPI_ACT_ROWSUM || a->c_dtype == MMVQA_BF16) && a->N >= 2 * VT_BN;
}

int vistok_launch(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  const bool fwd = a->epilogue == MMVQA_EPI_ACT_ROWSUM;
  if (a->K <= 64) return fwd ? launch_vt<0, 1, 3>(a, ep, st) : launch_vt<1, 1, 3>(a, ep, st);
  return fwd ? launch_vt<0, 2, 2>(a, ep, st) : launch_vt<1, 2, 2>(a, ep, st);
}

}  // namespace mmvqa
