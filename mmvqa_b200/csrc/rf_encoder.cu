// rf_encoder.cu -- the whole RealFormer encoder forward (models/realformer.py:30-51 x n_layers, mmbert.py:103-108) as
// ONE launch for small batches: a sample-stationary cluster kernel.
//
// Why: at the flagship shape (B = 16, T = 28, M = 448 tokens) the per-operator path is a chain of 6 dependent launches
// per layer whose fixed cost (launch, set-up, first-tile latency, epilogue) exceeds their main loops (DESIGN.md
// section 8).  Samples are independent through the encoder, so the chain can stay on chip:
//   * one 8-CTA cluster owns a group of samples (<= 64 token rows) for all layers; CTA h of the cluster owns head h of
//     the attention, output features [96h, 96h+96) of proj / FF2 and hidden features [384h, 384h+384) of FF1;
//   * every GEMM is computed TRANSPOSED on the tensor cores: D^T[features (UMMA M = 128), tokens (UMMA N = 64)] =
//     W[features, K] . X^T -- the weight slice is the A operand (streamed once per layer through a TMA ring, 48 KB
//     boxes {64 k, rows, 3-4 k-chunks} with 128-byte swizzle), the activations are the B operand, the fp32 accumulator
//     lives in TMEM (one lane = one feature, one column = one token);
//   * activations cross CTAs only through global memory that the backward pass needs anyway (attention output, x1,
//     SERF(h)): a CTA stores its slice, the cluster synchronises through remote mbarrier arrivals, and the full rows
//     come back as the next B operand with ONE TMA box (4-D map {64, rows, k-chunks, layer});
//   * LayerNorm row statistics are exchanged through distributed shared memory (per-CTA two-pass statistics merged
//     with Chan's formula);
//   * the kqv projection and the T x T attention run on mma.sync fragments (T <= 32: too small for a 128-row UMMA),
//     the RealFormer residual scores are read from / written to the fp32 score tensor of the previous / this layer.
// Warp roles (384 threads): warp 0 = weight-ring producer, warp 1 = activation producer (gathers, Wkqv), warp 2 = MMA
// issuer, warp 3 = TMEM allocator, warps 4..11 = attention + epilogues (warp w reads TMEM lanes 32 (w % 4) ...).
// Measured ingest behind the tile sizes (tools/ubench/stream_bench2.cu): a TMA op costs ~0.3 us + bytes / 180 GB/s per
// CTA whatever the ring depth, so the boxes are as large as shared memory allows.
#include "rf_cluster.cuh"

namespace mmvqa {

constexpr int RF_H = 768, RF_HEADS = RFC_HEADS, RF_D = 96, RF_F = 3072, RF_FS = 384;
constexpr int RF_N = 64;                  // UMMA N: token rows of one cluster (padded)
constexpr int RF_TP = 32;                 // attention tile (T <= 32)
constexpr int RF_MAXL = 16;
constexpr int RF_THREADS = 384;
constexpr int RF_A_STAGE = 49152, RF_NA = 2;
constexpr int RF_KC_H = RF_H / 64;        // 12 k-chunks of 64 over hidden
constexpr int RF_KC_F = RF_F / 64;        // 48 over the FF width
// shared-memory map (bytes)
constexpr int RF_A_OFF = 0;
constexpr int RF_B_OFF = RF_NA * RF_A_STAGE;               // 96 KB region: resident B operand / FF2 chunks / attention
constexpr int RF_B_BYTES = RF_N * RF_H * 2;                // 98304
constexpr int RF_WK_BYTES = 3 * RF_D * RF_D * 2;           // 55296: Wkqv [288][96]
constexpr int RF_LDN = RF_D + 8, RF_LDT = RF_TP + 8;       // padded rows of the mma.sync tiles
constexpr int RF_QKV_BYTES = (2 * RF_TP * RF_LDN + RF_D * RF_LDT) * 2;   // Q, K, V^T of one sample: 20992
constexpr int RF_YS_OFF = RF_B_OFF + RF_WK_BYTES;           // [64][104] bf16 LayerNorm input rows (aliases the Q/K/V^T scratch)
constexpr int RF_XS_OFF = RF_B_OFF + RF_B_BYTES;           // [64][104] bf16: this head's slice of the layer input
constexpr int RF_XS_BYTES = RF_N * RF_LDN * 2;
constexpr int RF_X1S_OFF = RF_XS_OFF + RF_XS_BYTES;        // [64][104] bf16: this head's slice of x1 (residual of the FF block)
constexpr int RF_STAT_OFF = RF_X1S_OFF + RF_XS_BYTES;      // [2 (ln)][8 (src)][64][2] floats
constexpr int RF_BAR_OFF = RF_STAT_OFF + 8192;
constexpr int RF_SMEM = RF_BAR_OFF + 256;
static_assert(RF_WK_BYTES + 2 * RF_QKV_BYTES <= RF_B_BYTES, "attention scratch must fit the B region");
static_assert(RF_SMEM <= 227 * 1024, "shared memory budget");
// TMEM columns
constexpr int RF_TM_PROJ = 0, RF_TM_FF1 = 64, RF_TM_FF2 = 256, RF_TM_COLS = 512;

struct RfEncParams {
  CUtensorMap wp[RF_MAXL], w1[RF_MAXL], w2[RF_MAXL];   // weights: 3-D {64, rows, k-chunks}
  CUtensorMap tm_att, tm_x1, tm_hact;                  // gathers: 4-D {64, M, k-chunks, layer}
  const bf16* wkqv[RF_MAXL];
  const void* wp_raw[RF_MAXL];
  const void* w1_raw[RF_MAXL];
  const void* w2_raw[RF_MAXL];
  const float* b1[RF_MAXL];
  const float* b2[RF_MAXL];
  const float* g1[RF_MAXL];
  const float* be1[RF_MAXL];
  const float* g2[RF_MAXL];
  const float* be2[RF_MAXL];
  const bf16* x0;   // [M, H]        encoder input
  bf16* xout;       // [L, M, H]     output of layer l (= input of layer l + 1)
  bf16* kqv;        // [L, M*8, 288]
  float* scores;    // [L, B, 8, T, T]
  bf16* att;        // [L, M, H]
  bf16* y1;         // [L, M, H]     x + dropout(proj(att))          (input of LN1)
  bf16* x1;         // [L, M, H]
  bf16* hpre;       // [L, M, F]
  bf16* hact;       // [L, M, F]
  bf16* y2;         // [L, M, H]     x1 + dropout(ff)                (input of LN2)
  float* mean1; float* rstd1; float* mean2; float* rstd2;   // [L, M]
  const float* prev0;   // optional [B, 8, T, T]
  const float* mask;    // optional [B, T]
  int B, T, L, spc, M;
  float p1, p2, eps;
  unsigned long long seed;
  const unsigned long long* seed_ctr;
  long long* trace;     // optional [4 roles][L][16] clock64 stamps of CTA 0 (tools/rf_encoder_check.py --trace)
};

// ---------------------------------------------------------------------------------------------------------------
// LayerNorm over the 768 features of each of the 64 token columns; this CTA holds 96 of them per token in
// Ys[token][96] (bf16, already rounded: the statistics are those of the stored tensor, as in add_ln_fwd_*).
// Thread (token c = ctid / 4, part = ctid % 4) owns 24 features: two-pass statistics inside the CTA, (sum, M2) of the
// 8 CTAs exchanged through distributed shared memory and merged with Chan's formula, then y_out = (y - mean) * rstd *
// gamma + beta goes to shared memory (next operand / residual) and, with the LayerNorm input itself, to global memory
// with 128-bit stores.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_layernorm(uint8_t* smem, int ctid, uint32_t rank, int ln_which, uint32_t cl_bar,
                                                  uint32_t parity, float eps, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, int nrows, int64_t row0,
                                                  bf16* __restrict__ ysum_out, bf16* __restrict__ y_out,
                                                  bf16* out_smem, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const bf16* Ys = reinterpret_cast<const bf16*>(smem + RF_YS_OFF);
  float* stat = reinterpret_cast<float*>(smem + RF_STAT_OFF) + ln_which * (RF_HEADS * RF_N * 2);   // [8][64][2]
  const int c = ctid >> 2, part = ctid & 3;
  const int f0 = 24 * part;
  compute_sync();                                       // Ys complete
  float v[24];
  uint4 raw[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    raw[i] = *reinterpret_cast<const uint4*>(Ys + c * RF_LDN + f0 + 8 * i);
    unpack8(raw[i], v + 8 * i);
  }
  float lsum = 0.0f;
#pragma unroll
  for (int i = 0; i < 24; ++i) lsum += v[i];
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 1);
  lsum += __shfl_xor_sync(0xffffffffu, lsum, 2);
  const float lmean = lsum * (1.0f / (float)RF_D);
  float lm2 = 0.0f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const float d = v[i] - lmean;
    lm2 += d * d;
  }
  lm2 += __shfl_xor_sync(0xffffffffu, lm2, 1);
  lm2 += __shfl_xor_sync(0xffffffffu, lm2, 2);
  if (part == 0) {
    const uint32_t local = smem_u32(stat + ((int)rank * RF_N + c) * 2);
#pragma unroll
    for (uint32_t r = 0; r < RF_HEADS; ++r) st_cluster_f32x2(map_to_cta(local, r), lsum, lm2);
  }
  const int64_t goff = (row0 + c) * RF_H + RF_D * (int)rank + f0;
  if (c < nrows) {
#pragma unroll
    for (int i = 0; i < 3; ++i) *reinterpret_cast<uint4*>(ysum_out + goff + 8 * i) = raw[i];
  }
  compute_sync();                                       // every publishing thread has issued its remote stores
  cluster_publish(cl_bar, ctid);
  mbar_wait_cluster(cl_bar, parity);
  float tot = 0.0f;
#pragma unroll
  for (int s = 0; s < RF_HEADS; ++s) tot += stat[(s * RF_N + c) * 2];
  const float mean = tot * (1.0f / (float)RF_H);
  float m2 = 0.0f;
#pragma unroll
  for (int s = 0; s < RF_HEADS; ++s) {
    const float ds = stat[(s * RF_N + c) * 2] * (1.0f / (float)RF_D) - mean;
    m2 += stat[(s * RF_N + c) * 2 + 1] + (float)RF_D * ds * ds;
  }
  const float rstd = rsqrtf(m2 * (1.0f / (float)RF_H) + eps);
  const float* gm = gamma + RF_D * (int)rank + f0;
  const float* bt = beta + RF_D * (int)rank + f0;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    uint4 o;
    uint32_t* ow = &o.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = 8 * i + 2 * k;
      const float a0 = (v[e] - mean) * rstd * __ldg(gm + e) + __ldg(bt + e);
      const float a1 = (v[e + 1] - mean) * rstd * __ldg(gm + e + 1) + __ldg(bt + e + 1);
      ow[k] = c < nrows ? pack2(a0, a1) : 0u;
    }
    *reinterpret_cast<uint4*>(out_smem + c * RF_LDN + f0 + 8 * i) = o;
    if (c < nrows) *reinterpret_cast<uint4*>(y_out + goff + 8 * i) = o;
  }
  if (rank == 0 && part == 0 && c < nrows) {
    mean_out[row0 + c] = mean;
    rstd_out[row0 + c] = rstd;
  }
}

#define RF_TRACE(role, slot)                                                                  \
  do {                                                                                        \
    if (p.trace && blockIdx.x == 0) p.trace[((role) * RF_MAXL + l) * 16 + (slot)] = clock64(); \
  } while (0)

// ---------------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(RF_HEADS, 1, 1) __launch_bounds__(RF_THREADS, 1)
    rf_encoder_fwd_kernel(const __grid_constant__ RfEncParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t h = cluster_ctarank();                 // head / feature-slice index
  const int grp = (int)cluster_id_x();                  // sample group of this cluster
  const int T = p.T, L = p.L, M = p.M;
  const int s0 = grp * p.spc;
  const int nsamp = min(p.spc, p.B - s0);
  const int row0 = s0 * T;                              // first token row of the group in [M, *]
  const int nrows = nsamp * T;
  // barriers
  const uint32_t bars = sbase + RF_BAR_OFF;
  const uint32_t a_full = bars, a_empty = bars + 16, b_full = bars + 32, b_empty = bars + 48;
  const uint32_t bres_full = bars + 64, bres_free = bars + 72, wkqv_full = bars + 80;
  const uint32_t tmem_full = bars + 88;                 // [5]: proj, ff1 x 3, ff2
  const uint32_t cl_s1 = bars + 128, cl_e1 = bars + 136, cl_s2 = bars + 144, cl_s3 = bars + 152, cl_e2 = bars + 160;
  const uint32_t tmem_ptr_addr = bars + 192;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem + RF_BAR_OFF + 192);
  volatile int* cur_layer = reinterpret_cast<volatile int*>(smem + RF_BAR_OFF + 200);

  if ((sbase & 1023u) != 0) __trap();                   // swizzled tiles need a 1024-byte aligned base
  if (threadIdx.x == 0) {
    *cur_layer = -1;
    for (int i = 0; i < 2; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
      mbar_init(b_full + 8 * i, 1);
      mbar_init(b_empty + 8 * i, 1);
    }
    mbar_init(bres_full, 1);
    mbar_init(bres_free, 1);
    mbar_init(wkqv_full, 1);
    for (int i = 0; i < 5; ++i) mbar_init(tmem_full + 8 * i, 1);
    mbar_init(cl_s1, RF_HEADS);
    mbar_init(cl_e1, RF_HEADS);
    mbar_init(cl_s2, RF_HEADS);
    mbar_init(cl_s3, RF_HEADS);
    mbar_init(cl_e2, RF_HEADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) tmem_alloc(tmem_ptr_addr, RF_TM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  pdl_wait();
  pdl_trigger();
  cluster_barrier_all();                                // every CTA's barriers exist before any remote arrival

  if (warp == 0) {
    // =========================== weight-ring producer ===========================
    if (lane == 0) {
      uint32_t it = 0;
      auto slot = [&](uint32_t& full, uint32_t& dst) {
        const uint32_t s = it % RF_NA, ph = (it / RF_NA) & 1u;
        mbar_wait(a_empty + 8 * s, ph ^ 1u);
        full = a_full + 8 * s;
        dst = sbase + RF_A_OFF + s * RF_A_STAGE;
        mbar_expect_tx(full, RF_A_STAGE);
        ++it;
      };
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        uint32_t full, dst;
        RF_TRACE(0, 0);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {                   // proj: rows [96h, 96h+96), 4 k-chunks per box
          slot(full, dst);
          tma_load_3d(dst, &p.wp[l], full, 0, 96 * (int)h, 4 * c);
        }
        RF_TRACE(0, 1);
#pragma unroll 1
        for (int m = 0; m < 3; ++m)                     // FF1: rows [384h + 128m, +128), 3 k-chunks per box
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            slot(full, dst);
            tma_load_3d(dst, &p.w1[l], full, 0, RF_FS * (int)h + 128 * m, 3 * c);
          }
        RF_TRACE(0, 2);
#pragma unroll 1
        for (int c = 0; c < 12; ++c) {                  // FF2: rows [96h, 96h+96), K = 3072
          slot(full, dst);
          tma_load_3d(dst, &p.w2[l], full, 0, 96 * (int)h, 4 * c);
        }
        RF_TRACE(0, 3);
      }
    }
  } else if (warp == 1) {
    // =========================== activation producer ===========================
    if (lane == 0) {
      uint32_t nfree = 0;                               // completed phases of bres_free consumed so far
      uint32_t nchunk = 0;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        const uint32_t lp = (uint32_t)l & 1u;
        // Wkqv of this layer: the B region is free once the FF2 MMAs of the previous layer have retired
        RF_TRACE(1, 0);
        if (l > 0) { mbar_wait(bres_free, nfree & 1u); ++nfree; }
        RF_TRACE(1, 1);
        mbar_expect_tx(wkqv_full, RF_WK_BYTES);
        bulk_load_1d(sbase + RF_B_OFF, p.wkqv[l], RF_WK_BYTES, wkqv_full);
        // attention output of the whole group: every CTA has stored its head
        mbar_wait_cluster(cl_s1, lp);
        RF_TRACE(1, 2);
        fence_proxy_async();
        mbar_expect_tx(bres_full, RF_B_BYTES);
        tma_load_4d(sbase + RF_B_OFF, &p.tm_att, bres_full, 0, row0, 0, l);
        // x1 of the whole group (after the proj MMAs have finished reading the attention rows)
        mbar_wait_cluster(cl_s2, lp);
        RF_TRACE(1, 3);
        mbar_wait(bres_free, nfree & 1u); ++nfree;
        RF_TRACE(1, 4);
        fence_proxy_async();
        mbar_expect_tx(bres_full, RF_B_BYTES);
        tma_load_4d(sbase + RF_B_OFF, &p.tm_x1, bres_full, 0, row0, 0, l);
        // SERF(h) of the whole group in 8 chunks of 6 k-chunks (48 KB), two slots
        mbar_wait_cluster(cl_s3, lp);
        RF_TRACE(1, 5);
        mbar_wait(bres_free, nfree & 1u); ++nfree;
        RF_TRACE(1, 6);
        fence_proxy_async();
#pragma unroll 1
        for (int c = 0; c < 8; ++c, ++nchunk) {
          const uint32_t s = nchunk & 1u, ph = (nchunk >> 1) & 1u;
          mbar_wait(b_empty + 8 * s, ph ^ 1u);
          mbar_expect_tx(b_full + 8 * s, RF_B_BYTES / 2);
          tma_load_4d(sbase + RF_B_OFF + s * (RF_B_BYTES / 2), &p.tm_hact, b_full + 8 * s, 0, row0, 6 * c, l);
        }
        RF_TRACE(1, 7);
      }
    }
  } else if (warp == 2) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      // D = f32, A = B = bf16, both K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(RF_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      uint32_t it = 0, nchunk = 0, nres = 0;
      uint32_t a_base = 0;
      auto a_acquire = [&]() {
        const uint32_t s = it % RF_NA, ph = (it / RF_NA) & 1u;
        mbar_wait(a_full + 8 * s, ph);
        tc_fence_after();
        a_base = sbase + RF_A_OFF + s * RF_A_STAGE;
      };
      auto a_release = [&]() {
        umma_commit(a_empty + 8 * (it % RF_NA));
        ++it;
      };
      auto mma_kc = [&](uint32_t a_kc, uint32_t b_kc, uint32_t d_tmem, bool first) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16(d_tmem, make_sdesc(a_kc + j * 32, 16, 1024), make_sdesc(b_kc + j * 32, 16, 1024), idesc,
                    (first && j == 0) ? 0u : 1u);
      };
      const uint32_t bres = sbase + RF_B_OFF;
#pragma unroll 1
      for (int l = 0; l < L; ++l) {
        // ---- proj: K = 768, A boxes of 96 rows x 4 k-chunks (12288 B per chunk), B resident (8192 B per chunk)
        RF_TRACE(2, 0);
        mbar_wait(bres_full, nres & 1u); ++nres;
        tc_fence_after();
        RF_TRACE(2, 1);
#pragma unroll 1
        for (int c = 0; c < 3; ++c) {
          a_acquire();
#pragma unroll 1
          for (int k = 0; k < 4; ++k) mma_kc(a_base + k * 12288, bres + (4 * c + k) * 8192, tmem + RF_TM_PROJ, c == 0 && k == 0);
          a_release();
        }
        umma_commit(tmem_full + 0);
        umma_commit(bres_free);
        RF_TRACE(2, 2);
        // ---- FF1: 3 tiles of 128 hidden features, A boxes of 128 rows x 3 k-chunks (16384 B per chunk)
        mbar_wait(bres_full, nres & 1u); ++nres;
        tc_fence_after();
        RF_TRACE(2, 3);
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            a_acquire();
#pragma unroll 1
            for (int k = 0; k < 3; ++k)
              mma_kc(a_base + k * 16384, bres + (3 * c + k) * 8192, tmem + RF_TM_FF1 + m * RF_N, c == 0 && k == 0);
            a_release();
          }
          umma_commit(tmem_full + 8 * (1 + m));
          RF_TRACE(2, 4 + m);
        }
        umma_commit(bres_free);
        // ---- FF2: K = 3072; A boxes of 96 rows x 4 k-chunks, B chunks of 6 k-chunks in two slots
        uint32_t b_base = 0;
#pragma unroll 1
        for (int kc = 0; kc < RF_KC_F; ++kc) {
          if (kc % 6 == 0) {
            const uint32_t s = nchunk & 1u, ph = (nchunk >> 1) & 1u;
            mbar_wait(b_full + 8 * s, ph);
            tc_fence_after();
            b_base = bres + s * (RF_B_BYTES / 2);
            if (kc == 0) RF_TRACE(2, 7);
          }
          if (kc % 4 == 0) a_acquire();
          mma_kc(a_base + (kc % 4) * 12288, b_base + (kc % 6) * 8192, tmem + RF_TM_FF2, kc == 0);
          if (kc % 4 == 3) a_release();
          if (kc % 6 == 5) {
            umma_commit(b_empty + 8 * (nchunk & 1u));
            ++nchunk;
          }
        }
        umma_commit(tmem_full + 8 * 4);
        umma_commit(bres_free);
        RF_TRACE(2, 8);
      }
    }
  } else if (warp == 3) {
    // =========================== L2 prefetcher ===========================
    // The weight boxes are first touches of HBM (the fp32 masters and Adam state stream 1.5 GB through L2 every step), and
    // the ring holds only 96 KB in flight: with HBM latency under every TMA op the ring is latency bound.  This warp
    // touches the NEXT layer's weight slices of this CTA with prefetch.global.L2 while the current layer computes (the
    // clusters share the lines: cluster g takes lines g, g + nclusters, ...), so the TMA boxes find them in L2.
    const int ncl = (int)gridDim.x / RF_HEADS;
#pragma unroll 1
    for (int l = 0; l + 1 < L; ++l) {
      while (*cur_layer < l) __nanosleep(200);          // layer l has started (counter written by the compute warps)
      const char* base[3] = {reinterpret_cast<const char*>(p.wp_raw[l + 1]) + (size_t)RF_D * h * RF_H * 2,
                             reinterpret_cast<const char*>(p.w1_raw[l + 1]) + (size_t)RF_FS * h * RF_H * 2,
                             reinterpret_cast<const char*>(p.w2_raw[l + 1]) + (size_t)RF_D * h * RF_F * 2};
      const int nline[3] = {RF_D * RF_H * 2 / 128, RF_FS * RF_H * 2 / 128, RF_D * RF_F * 2 / 128};
#pragma unroll 1
      for (int w = 0; w < 3; ++w)
#pragma unroll 1
        for (int i = grp + ncl * lane; i < nline[w]; i += ncl * 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base[w] + (size_t)i * 128));
      if (grp == 0 && h == 0)
        for (int i = lane; i < RF_WK_BYTES / 128; i += 32)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.wkqv[l + 1]) + (size_t)i * 128));
    }
  } else if (warp >= 4) {
    // =========================== attention + epilogues (256 threads) ===========================
    const int cw = warp - 4;                            // 0..7
    const int q = cw & 3, ch = cw >> 2;                 // TMEM lane quadrant, column half
    const int ctid = threadIdx.x - 128;                 // 0..255
    const int g = lane >> 2, t4 = lane & 3;
    bf16* Xs = reinterpret_cast<bf16*>(smem + RF_XS_OFF);                       // [64][104]
    const bf16* Ws = reinterpret_cast<const bf16*>(smem + RF_B_OFF);            // [288][96] (TMA, dense)
    uint8_t* qkv_base = smem + RF_B_OFF + RF_WK_BYTES;                          // per sample: Q, K [32][104], V^T [96][40]
    const int f_loc = 32 * q + lane;                    // feature inside this CTA's 96-wide slice (q < 3)
    const bool feat_ok = q < 3;
    const int f_glob = RF_D * (int)h + f_loc;
    const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
    const unsigned long long seed0 = seed_eff(p.seed, p.seed_ctr);
    const uint32_t thr1 = (uint32_t)(p.p1 * 4294967296.0), thr2 = (uint32_t)(p.p2 * 4294967296.0);
    const float keep1 = p.p1 > 0.0f ? 1.0f / (1.0f - p.p1) : 1.0f, keep2 = p.p2 > 0.0f ? 1.0f / (1.0f - p.p2) : 1.0f;
    const float inv_sqrt_d = 1.0f / sqrtf((float)RF_D);
    const int64_t MH = (int64_t)M * RF_H, MF = (int64_t)M * RF_F;
    const int64_t score_layer = (int64_t)p.B * RF_HEADS * T * T;

    // layer-0 input slice -> Xs (rows beyond the group are zero)
    for (int i = ctid; i < RF_N * (RF_D / 8); i += 256) {
      const int r = i / (RF_D / 8), c8 = i % (RF_D / 8);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < nrows) v = *reinterpret_cast<const uint4*>(p.x0 + (int64_t)(row0 + r) * RF_H + RF_D * h + c8 * 8);
      *reinterpret_cast<uint4*>(Xs + r * RF_LDN + c8 * 8) = v;
    }
    compute_sync();
    bf16* X1s = reinterpret_cast<bf16*>(smem + RF_X1S_OFF);                     // [64][104]
    bf16* Ys = reinterpret_cast<bf16*>(smem + RF_YS_OFF);                       // [64][104]

#pragma unroll 1
    for (int l = 0; l < L; ++l) {
      const uint32_t lp = (uint32_t)l & 1u;
      // ------------------------------------------------------------------ kqv + attention (mma.sync)
      // the FF2 accumulator of the previous layer has been read (end of the previous iteration), so the B region is
      // free: zero the Q / K / V^T tiles (padding rows must be finite zeros)
      {
        uint4* z = reinterpret_cast<uint4*>(qkv_base);
        for (int i = ctid; i < 2 * RF_QKV_BYTES / 16; i += 256) z[i] = make_uint4(0, 0, 0, 0);
      }
      if (ctid == 0) RF_TRACE(3, 0);
      if (ctid == 0) *cur_layer = l;
      mbar_wait(wkqv_full, lp);
      compute_sync();
      if (ctid == 0) RF_TRACE(3, 1);
      {
        // kqv[r, n] = sum_k Xs[r, k] Ws[n, k]: n-tiles of 8 over the 8 warps, B fragments hoisted over the 4 row strips
        bf16* kqv_out = p.kqv + (int64_t)l * M * RF_HEADS * 3 * RF_D;
        for (int nt = cw; nt < 3 * RF_D / 8; nt += 8) {
          uint32_t bfr[RF_D / 16][2];
#pragma unroll
          for (int kd = 0; kd < RF_D / 16; ++kd) load_b(bfr[kd], Ws, RF_D, nt * 8, kd * 16, g, t4);
          const int n = nt * 8 + 2 * t4;
          const int sec = n / RF_D, cc = n - sec * RF_D;          // 0: k, 1: q, 2: v (realformer.py:33)
#pragma unroll
          for (int strip = 0; strip < RF_N / 16; ++strip) {
            if (strip * 16 >= nrows) break;
            float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int kd = 0; kd < RF_D / 16; ++kd) {
              uint32_t a[4];
              load_a(a, Xs, RF_LDN, strip * 16, kd * 16, g, t4);
              mma_bf16_16816(c, a, bfr[kd]);
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int r = strip * 16 + g + 8 * half;
              if (r >= nrows) continue;
              const uint32_t val = pack2(c[2 * half], c[2 * half + 1]);
              const int s = r >= T ? 1 : 0, tq = r - s * T;      // at most two samples per cluster
              bf16* Qs = reinterpret_cast<bf16*>(qkv_base + s * RF_QKV_BYTES);
              bf16* Ks = Qs + RF_TP * RF_LDN;
              bf16* Vt = Ks + RF_TP * RF_LDN;
              if (sec == 2) {
                const __nv_bfloat162 pv = *reinterpret_cast<const __nv_bfloat162*>(&val);
                Vt[cc * RF_LDT + tq] = pv.x;
                Vt[(cc + 1) * RF_LDT + tq] = pv.y;
              } else {
                *reinterpret_cast<uint32_t*>((sec == 0 ? Ks : Qs) + tq * RF_LDN + cc) = val;
              }
              *reinterpret_cast<uint32_t*>(kqv_out + ((int64_t)(row0 + r) * RF_HEADS + h) * (3 * RF_D) + n) = val;
            }
          }
        }
      }
      compute_sync();
      if (ctid == 0) RF_TRACE(3, 2);
      {
        // attention of sample cw / 2, query strip cw % 2 (T <= 32: two 16-row strips per sample)
        const int s = cw >> 1, r0 = (cw & 1) * 16;
        if (s < nsamp && r0 < T) {
          const bf16* Qs = reinterpret_cast<const bf16*>(qkv_base + s * RF_QKV_BYTES);
          const bf16* Ks = Qs + RF_TP * RF_LDN;
          const bf16* Vt = Ks + RF_TP * RF_LDN;
          const int b = s0 + s;
          const int iA = r0 + g, iB = r0 + g + 8;
          float sc[RF_TP / 8][4];
#pragma unroll
          for (int nt = 0; nt < RF_TP / 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.0f;
#pragma unroll
          for (int kd = 0; kd < RF_D / 16; ++kd) {
            uint32_t a[4];
            load_a(a, Qs, RF_LDN, r0, kd * 16, g, t4);
#pragma unroll
            for (int nt = 0; nt < RF_TP / 8; ++nt) {
              uint32_t bb[2];
              load_b(bb, Ks, RF_LDN, nt * 8, kd * 16, g, t4);
              mma_bf16_16816(sc[nt], a, bb);
            }
          }
          const float* prev = l == 0 ? p.prev0 : p.scores + (int64_t)(l - 1) * score_layer;
          float* sout = p.scores + (int64_t)l * score_layer;
          const int64_t sb = ((int64_t)b * RF_HEADS + h) * T * T;
          float qoffA = 0.0f, qoffB = 0.0f;
          if (p.mask) {
            if (iA < T) qoffA = -10000.0f * (1.0f - __ldg(p.mask + b * T + iA));
            if (iB < T) qoffB = -10000.0f * (1.0f - __ldg(p.mask + b * T + iB));
          }
          float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < RF_TP / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int i = (e < 2) ? iA : iB;
              const int j = nt * 8 + 2 * t4 + (e & 1);
              float v = -INFINITY;
              if (i < T && j < T) {
                v = sc[nt][e] * inv_sqrt_d;
                if (prev) v += prev[sb + (int64_t)i * T + j];
                v += (e < 2) ? qoffA : qoffB;
                sout[sb + (int64_t)i * T + j] = v;
              }
              sc[nt][e] = v;
              if (e < 2) mxA = fmaxf(mxA, v); else mxB = fmaxf(mxB, v);
            }
          }
          mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1));
          mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
          mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1));
          mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
          float sumA = 0.0f, sumB = 0.0f;
#pragma unroll
          for (int nt = 0; nt < RF_TP / 8; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float m = (e < 2) ? mxA : mxB;
              const float pr = (sc[nt][e] == -INFINITY) ? 0.0f : expf(sc[nt][e] - m);
              sc[nt][e] = pr;
              if (e < 2) sumA += pr; else sumB += pr;
            }
          }
          sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
          sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
          sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
          sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
          const float invA = sumA > 0.0f ? 1.0f / sumA : 0.0f, invB = sumB > 0.0f ? 1.0f / sumB : 0.0f;
#pragma unroll
          for (int nt = 0; nt < RF_TP / 8; ++nt) {
            sc[nt][0] *= invA; sc[nt][1] *= invA; sc[nt][2] *= invB; sc[nt][3] *= invB;
          }
          float o[RF_D / 8][4];
#pragma unroll
          for (int nd = 0; nd < RF_D / 8; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.0f;
#pragma unroll
          for (int ks = 0; ks < RF_TP / 16; ++ks) {
            uint32_t a[4];
            a[0] = pack2(sc[2 * ks][0], sc[2 * ks][1]);
            a[1] = pack2(sc[2 * ks][2], sc[2 * ks][3]);
            a[2] = pack2(sc[2 * ks + 1][0], sc[2 * ks + 1][1]);
            a[3] = pack2(sc[2 * ks + 1][2], sc[2 * ks + 1][3]);
#pragma unroll
            for (int nd = 0; nd < RF_D / 8; ++nd) {
              uint32_t bb[2];
              load_b(bb, Vt, RF_LDT, nd * 8, ks * 16, g, t4);
              mma_bf16_16816(o[nd], a, bb);
            }
          }
          bf16* aout = p.att + (int64_t)l * MH;
#pragma unroll
          for (int nd = 0; nd < RF_D / 8; ++nd) {
            const int cidx = RF_D * (int)h + nd * 8 + 2 * t4;
            if (iA < T) *reinterpret_cast<uint32_t*>(aout + ((int64_t)b * T + iA) * RF_H + cidx) = pack2(o[nd][0], o[nd][1]);
            if (iB < T) *reinterpret_cast<uint32_t*>(aout + ((int64_t)b * T + iB) * RF_H + cidx) = pack2(o[nd][2], o[nd][3]);
          }
        }
      }
      // publish: the attention rows of this head are in global memory, the B region may be overwritten by the gather
      fence_proxy_async();
      compute_sync();
      cluster_publish(cl_s1, ctid);
      if (ctid == 0) RF_TRACE(3, 3);

      // ------------------------------------------------------------------ proj epilogue: dropout, + residual -> Ys; LN1
      mbar_wait(tmem_full + 0, lp);
      tc_fence_after();
      if (ctid == 0) RF_TRACE(3, 4);
      if (feat_ok) {
        const unsigned long long sd = seed0 + 2ull * (unsigned long long)l;
#pragma unroll 1
        for (int cb = 0; cb < 32; cb += 8) {
          uint32_t r[8];
          __syncwarp();
          tmem_ld8(tmem_lane + RF_TM_PROJ + ch * 32 + cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = ch * 32 + cb + k;
            float a = __uint_as_float(r[k]);
            if (p.p1 > 0.0f) a = hash32(sd, (uint64_t)(row0 + c) * RF_H + (uint64_t)f_glob) >= thr1 ? a * keep1 : 0.0f;
            a += __bfloat162float(Xs[c * RF_LDN + f_loc]);
            Ys[c * RF_LDN + f_loc] = __float2bfloat16_rn(a);
          }
        }
      }
      if (ctid == 0) RF_TRACE(3, 5);
      cluster_layernorm(smem, ctid, h, 0, cl_e1, lp, p.eps, p.g1[l], p.be1[l], nrows, row0, p.y1 + (int64_t)l * MH,
                        p.x1 + (int64_t)l * MH, X1s, p.mean1 + (int64_t)l * M, p.rstd1 + (int64_t)l * M);
      if (ctid == 0) RF_TRACE(3, 6);
      fence_proxy_async();
      compute_sync();
      cluster_publish(cl_s2, ctid);
      if (ctid == 0) RF_TRACE(3, 7);

      // ------------------------------------------------------------------ FF1 epilogue: + bias, SERF
      {
        bf16* hpo = p.hpre + (int64_t)l * MF;
        bf16* hao = p.hact + (int64_t)l * MF;
#pragma unroll 1
        for (int m = 0; m < 3; ++m) {
          mbar_wait(tmem_full + 8 * (1 + m), lp);
          tc_fence_after();
          if (ctid == 0) RF_TRACE(3, 8 + m);
          const int fh = RF_FS * (int)h + 128 * m + 32 * q + lane;
          const float bias = __ldg(p.b1[l] + fh);
#pragma unroll 1
          for (int cb = 0; cb < 32 && ch * 32 + cb < nrows; cb += 8) {
            uint32_t r[8];
            __syncwarp();
            tmem_ld8(tmem_lane + RF_TM_FF1 + m * RF_N + ch * 32 + cb, r);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const int c = ch * 32 + cb + k;
              const float a = __uint_as_float(r[k]) + bias;
              if (c < nrows) {
                const int64_t off = (int64_t)(row0 + c) * RF_F + fh;
                hpo[off] = __float2bfloat16_rn(a);
                hao[off] = __float2bfloat16_rn(serf_fast(a));
              }
            }
          }
        }
      }
      fence_proxy_async();
      compute_sync();
      cluster_publish(cl_s3, ctid);
      if (ctid == 0) RF_TRACE(3, 11);

      // ------------------------------------------------------------------ FF2 epilogue: + bias, dropout, + x1 -> Ys; LN2
      mbar_wait(tmem_full + 8 * 4, lp);
      tc_fence_after();
      if (ctid == 0) RF_TRACE(3, 12);
      if (feat_ok) {
        const unsigned long long sd = seed0 + 2ull * (unsigned long long)l + 1ull;
        const float bias = __ldg(p.b2[l] + f_glob);
#pragma unroll 1
        for (int cb = 0; cb < 32; cb += 8) {
          uint32_t r[8];
          __syncwarp();
          tmem_ld8(tmem_lane + RF_TM_FF2 + ch * 32 + cb, r);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = ch * 32 + cb + k;
            float a = __uint_as_float(r[k]) + bias;
            if (p.p2 > 0.0f) a = hash32(sd, (uint64_t)(row0 + c) * RF_H + (uint64_t)f_glob) >= thr2 ? a * keep2 : 0.0f;
            a += __bfloat162float(X1s[c * RF_LDN + f_loc]);
            Ys[c * RF_LDN + f_loc] = __float2bfloat16_rn(a);
          }
        }
      }
      if (ctid == 0) RF_TRACE(3, 13);
      cluster_layernorm(smem, ctid, h, 1, cl_e2, lp, p.eps, p.g2[l], p.be2[l], nrows, row0, p.y2 + (int64_t)l * MH,
                        p.xout + (int64_t)l * MH, Xs, p.mean2 + (int64_t)l * M, p.rstd2 + (int64_t)l * M);
      if (ctid == 0) RF_TRACE(3, 14);
      compute_sync();                                   // Xs complete before the next layer's kqv reads it
      if (ctid == 0) RF_TRACE(3, 15);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier_all();                                // no CTA leaves while peers may still address its shared memory
  if (warp == 3) tmem_dealloc(tmem, RF_TM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFnRf)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnRf rf_get_encode() {
  static EncodeTiledFnRf fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFnRf>(sym);
  }
  return fn;
}
// weight [rows, K] bf16 row-major seen as {64, rows, K/64}
static int rf_weight_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows, int box_kc, const char* what) {
  EncodeTiledFnRf enc = rf_get_encode();
  if (!enc) return set_err(MMVQA_ERR_CUDA, "rf_encoder: cuTensorMapEncodeTiled unavailable");
  MMVQA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "rf_encoder: %s must be 16-byte aligned", what);
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_kc};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(MMVQA_ERR_CUDA, "rf_encoder: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return MMVQA_OK;
}
// activations [L, M, W] bf16 seen as {64, M, W/64, L}
static int rf_act_map(CUtensorMap* map, const void* base, int M, int W, int L, int box_kc, const char* what) {
  EncodeTiledFnRf enc = rf_get_encode();
  if (!enc) return set_err(MMVQA_ERR_CUDA, "rf_encoder: cuTensorMapEncodeTiled unavailable");
  MMVQA_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "rf_encoder: %s must be 16-byte aligned", what);
  cuuint64_t dims[4] = {64, (cuuint64_t)M, (cuuint64_t)(W / 64), (cuuint64_t)L};
  cuuint64_t strides[3] = {(cuuint64_t)W * 2, 128, (cuuint64_t)M * W * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)RF_N, (cuuint32_t)box_kc, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(MMVQA_ERR_CUDA, "rf_encoder: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
  return MMVQA_OK;
}

// samples per cluster: their token rows share the 64-column B operand; the attention scratch holds two Q/K/V^T sets
static int rf_samples_per_cluster(int T) { return RF_N / T < 2 ? RF_N / T : 2; }

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_rf_encoder_fwd_supported(int B, int T, int hidden, int heads, int ff, int n_layers) {
  if (hidden != RF_H || heads != RF_HEADS || ff != RF_F) return 0;
  if (T < 1 || T > RF_TP || n_layers < 1 || n_layers > RF_MAXL || B < 1) return 0;
  const int spc = rf_samples_per_cluster(T);
  const int groups = (B + spc - 1) / spc;
  return groups <= 14 ? 1 : 0;                          // all clusters must be co-resident (15 fit on a B200; keep one spare)
}

int mmvqa_rf_encoder_fwd(const mmvqa_rf_encoder_args* a, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(a != nullptr, "rf_encoder_fwd: null args");
  MMVQA_REQUIRE(mmvqa_rf_encoder_fwd_supported(a->B, a->T, a->hidden, a->heads, a->ff, a->n_layers),
                "rf_encoder_fwd: unsupported shape (needs hidden 768, 8 heads, ff 3072, T <= 32, <= 16 layers, <= 14 sample groups)");
  MMVQA_REQUIRE(a->x0 && a->xout && a->kqv && a->scores && a->att && a->y1 && a->x1 && a->hpre && a->hact && a->y2 && a->mean1 && a->rstd1 &&
                    a->mean2 && a->rstd2, "rf_encoder_fwd: null buffer");
  MMVQA_REQUIRE(a->dropout_p1 >= 0.0f && a->dropout_p1 < 1.0f && a->dropout_p2 >= 0.0f && a->dropout_p2 < 1.0f,
                "rf_encoder_fwd: dropout must be in [0,1)");
  MMVQA_REQUIRE((reinterpret_cast<uintptr_t>(a->x0) & 15) == 0, "rf_encoder_fwd: x0 must be 16-byte aligned");
  static RfEncParams prm;                               // ~8 KB: keep it off the stack
  const int L = a->n_layers, M = a->B * a->T;
  int rc;
  for (int l = 0; l < L; ++l) {
    MMVQA_REQUIRE(a->wkqv[l] && a->wproj[l] && a->w1[l] && a->w2[l] && a->b1[l] && a->b2[l] && a->ln1_w[l] && a->ln1_b[l] &&
                      a->ln2_w[l] && a->ln2_b[l], "rf_encoder_fwd: null parameter pointer in layer %d", l);
    MMVQA_REQUIRE((reinterpret_cast<uintptr_t>(a->wkqv[l]) & 15) == 0, "rf_encoder_fwd: kqv weight must be 16-byte aligned");
    if ((rc = rf_weight_map(&prm.wp[l], a->wproj[l], RF_H, RF_H, RF_D, 4, "proj.weight"))) return rc;
    if ((rc = rf_weight_map(&prm.w1[l], a->w1[l], RF_F, RF_H, 128, 3, "ff.0.weight"))) return rc;
    if ((rc = rf_weight_map(&prm.w2[l], a->w2[l], RF_H, RF_F, RF_D, 4, "ff.2.weight"))) return rc;
    prm.wkqv[l] = reinterpret_cast<const bf16*>(a->wkqv[l]);
    prm.wp_raw[l] = a->wproj[l]; prm.w1_raw[l] = a->w1[l]; prm.w2_raw[l] = a->w2[l];
    prm.b1[l] = a->b1[l]; prm.b2[l] = a->b2[l];
    prm.g1[l] = a->ln1_w[l]; prm.be1[l] = a->ln1_b[l]; prm.g2[l] = a->ln2_w[l]; prm.be2[l] = a->ln2_b[l];
  }
  if ((rc = rf_act_map(&prm.tm_att, a->att, M, RF_H, L, RF_KC_H, "att"))) return rc;
  if ((rc = rf_act_map(&prm.tm_x1, a->x1, M, RF_H, L, RF_KC_H, "x1"))) return rc;
  if ((rc = rf_act_map(&prm.tm_hact, a->hact, M, RF_F, L, 6, "hact"))) return rc;
  RfEncParams& P = prm;
  P.x0 = reinterpret_cast<const bf16*>(a->x0); P.xout = reinterpret_cast<bf16*>(a->xout); P.kqv = reinterpret_cast<bf16*>(a->kqv); P.scores = a->scores;
  P.att = reinterpret_cast<bf16*>(a->att); P.x1 = reinterpret_cast<bf16*>(a->x1); P.hact = reinterpret_cast<bf16*>(a->hact);
  P.y1 = reinterpret_cast<bf16*>(a->y1); P.hpre = reinterpret_cast<bf16*>(a->hpre); P.y2 = reinterpret_cast<bf16*>(a->y2);
  P.mean1 = a->mean1; P.rstd1 = a->rstd1; P.mean2 = a->mean2; P.rstd2 = a->rstd2;
  P.prev0 = a->prev; P.mask = a->mask;
  P.B = a->B; P.T = a->T; P.L = L; P.M = M;
  P.spc = rf_samples_per_cluster(a->T);
  P.p1 = a->dropout_p1; P.p2 = a->dropout_p2; P.eps = a->eps;
  P.seed = a->dropout_seed; P.seed_ctr = g_seed_ctr;
  P.trace = reinterpret_cast<long long*>(a->trace);
  const int groups = (a->B + P.spc - 1) / P.spc;
  static bool attr_set = false;
  if (!attr_set) {
    MMVQA_CUDA(cudaFuncSetAttribute(rf_encoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RF_SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * RF_HEADS);
  cfg.blockDim = dim3(RF_THREADS);
  cfg.dynamicSmemBytes = RF_SMEM;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  MMVQA_CUDA(cudaLaunchKernelEx(&cfg, rf_encoder_fwd_kernel, P));
  MMVQA_LAUNCHED("rf_encoder_fwd");
  return MMVQA_OK;
}

}  // extern "C"
