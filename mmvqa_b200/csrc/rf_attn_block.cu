// rf_attn_block.cu -- backward of the attention block of ONE RealFormer layer as one cluster launch (bf16, small batches):
//     LN1 backward  ->  proj dgrad  ->  residual-attention backward  ->  kqv dgrad (+ residual)
// i.e. the backward of  x1 = ln1(x + dropout(proj(resmha(x))))  (models/realformer.py:30-45,49), the four dependent
// launches (layernorm_bwd_parts | gemm dgrad | rf_attn_bwd | gemm dgrad+residual) that sit on the critical path of every
// layer of the backward pass (59 of its 98 us at B = 16, T = 28; profiles/r02_timeline_summary.txt).
//
// Same decomposition as rf_encoder.cu: one 8-CTA cluster per group of samples (<= 64 token rows), CTA h owns head h and
// the hidden features [96h, 96h + 96).
//   A. LN1 backward on the CTA's 96-feature slice of the 64 token rows: g = round_bf16(sum of the fp32 split-K slabs of
//      the FF1 dgrad + the residual gradient); the two row sums over the 768 features are exchanged through distributed
//      shared memory; dgamma / dbeta go to global memory with one atomic per (CTA, feature); dy1 stays in shared memory
//      (residual of the block), dropout(dy1) goes to global memory (operand of the proj weight-gradient GEMM).
//   B. the cluster synchronises (remote mbarrier arrivals) and every CTA gathers the full dropout(dy1) rows with ONE TMA
//      box as the B operand of
//   C. the transposed proj dgrad on the tensor cores: D^T[in-feature (UMMA M = 128, 96 used), token (N = 64)] =
//      Wp[out, in]^T . dpr^T -- the weight slice is the MN-major A operand (chunked 4-D TMA boxes {64 in, 128 out, 2}),
//      fp32 accumulator in TMEM.
//   D. the accumulator (dO of this head) is unloaded straight into the mma.sync tiles of
//   E. the residual-attention backward of the group's samples (P recomputed from the stored scores, running RealFormer
//      score gradient in / out, dQ / dK / dV) and the kqv input gradient dx = dkqv . Wkqv + dy1 (attention_tc.cuh's
//      arithmetic, two 16-row strips per sample, all 8 warps in the kqv dgrad).
// The weight-gradient GEMMs of proj and kqv stay on the side branch: they read dropout(dy1) and dkqv from global memory.
#include "rf_cluster.cuh"

namespace mmvqa {

constexpr int AB_H = 768, AB_D = 96, AB_N = 64, AB_TP = 32, AB_THREADS = 384;
constexpr int AB_A_STAGE = 32768, AB_NA = 3;                 // Wp boxes {64 in, 128 out, 2 groups}: 128 k-rows per stage
constexpr int AB_A_OFF = 0;
constexpr int AB_B_OFF = AB_NA * AB_A_STAGE;                 // 96 KB: dropout(dy1) of the whole group, K-major chunks
constexpr int AB_B_BYTES = AB_N * AB_H * 2;
constexpr int AB_LDN = AB_D + 8, AB_LDT = AB_TP + 8, AB_LDG = 3 * AB_D + 8;
constexpr int AB_DY_OFF = AB_B_OFF + AB_B_BYTES;             // [64][104] bf16 dy1 slice
constexpr int AB_DY_BYTES = AB_N * AB_LDN * 2;
constexpr int AB_STAT_OFF = AB_DY_OFF + AB_DY_BYTES;         // [8 src][64][2] floats
constexpr int AB_GB_OFF = AB_STAT_OFF + 4096;                // [2][96] floats: dgamma / dbeta of this CTA
constexpr int AB_BAR_OFF = AB_GB_OFF + 768;
constexpr int AB_SMEM = AB_BAR_OFF + 128;
// attention scratch (aliases the ring and the B region once the MMAs have retired)
constexpr int AB_WS_BYTES = 3 * AB_D * AB_LDN * 2;           // Wkqv [288][104]
constexpr int AB_SAMPLE_BYTES = (2 * AB_TP * AB_LDN + 3 * AB_D * AB_LDT + 2 * AB_TP * AB_LDT + AB_TP * AB_LDG) * 2;
constexpr int AB_SAMPLE_STRIDE = (AB_SAMPLE_BYTES + 127) & ~127;
static_assert(AB_WS_BYTES + 2 * AB_SAMPLE_STRIDE <= AB_DY_OFF, "attention scratch must fit the ring + B region");
static_assert(AB_SMEM <= 227 * 1024, "shared memory budget");

struct RfAttnBwdParams {
  CUtensorMap tm_wp;      // Wp [768 out, 768 in] as {64, 768, 12, 1}, box {64, 128, 2, 1}
  CUtensorMap tm_dpr;     // dropout(dy1) [M, 768] as {64, M, 12, 1}, box {64, 64, 12, 1}
  const float* parts; int nparts; long long part_stride;
  const bf16* dres;
  const bf16* y1; const float* mean1; const float* rstd1; const float* g1;
  const bf16* wkqv; const bf16* kqv; const float* scores; const float* ds_in;
  bf16* dpr; bf16* dkqv; float* dprev; bf16* dxin; float* dg1; float* db1;
  int B, T, spc, M;
  float p1;
  unsigned long long seed;
  const unsigned long long* seed_ctr;
};

// global [Tn, d] bf16 tile -> natural [row][ldn] and / or transposed [col][ldt] shared-memory copies, `nthr` threads
__device__ __forceinline__ void ab_stage_tile(const bf16* __restrict__ src, int64_t row_stride, int Tn, int d, bf16* nat,
                                              int ldn, bf16* tr, int ldt, int tid, int nthr) {
  const int cpr = d / 8, total = Tn * cpr;
  for (int idx = tid; idx < total; idx += nthr) {
    const int row = idx / cpr, ch = idx - row * cpr;
    const uint4 v = *reinterpret_cast<const uint4*>(src + row * row_stride + ch * 8);
    if (nat) *reinterpret_cast<uint4*>(nat + row * ldn + ch * 8) = v;
    if (tr) {
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const unsigned short bits = (unsigned short)((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
        reinterpret_cast<unsigned short*>(tr)[(ch * 8 + e) * ldt + row] = bits;
      }
    }
  }
}

__global__ void __cluster_dims__(RFC_HEADS, 1, 1) __launch_bounds__(AB_THREADS, 1)
    rf_attn_block_bwd_kernel(const __grid_constant__ RfAttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t h = cluster_ctarank();
  const int grp = (int)cluster_id_x();
  const int T = p.T, M = p.M;
  const int s0 = grp * p.spc;
  const int nsamp = min(p.spc, p.B - s0);
  const int row0 = s0 * T;
  const int nrows = nsamp * T;
  const uint32_t bars = sbase + AB_BAR_OFF;
  const uint32_t a_full = bars, a_empty = bars + 24, b_full = bars + 48, tmem_full = bars + 56, cl_e = bars + 64, cl_s = bars + 72;
  const uint32_t tmem_ptr_addr = bars + 96;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem + AB_BAR_OFF + 96);
  // the 128-row UMMA tile starts at the 64-aligned input feature below 96h: the head's features sit at lanes
  // lane_off .. lane_off + 95 of the accumulator (0 for even heads, 32 for odd heads)
  const int m_base = (AB_D * (int)h) / 64 * 64;
  const int lane_off = AB_D * (int)h - m_base;

  if ((sbase & 1023u) != 0) __trap();
  if (threadIdx.x == 0) {
    for (int i = 0; i < AB_NA; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    mbar_init(b_full, 1);
    mbar_init(tmem_full, 1);
    mbar_init(cl_e, RFC_HEADS);
    mbar_init(cl_s, RFC_HEADS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) tmem_alloc(tmem_ptr_addr, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_gen;
  if (warp == 0 && lane == 0) {
    // the weight boxes do not depend on the previous kernel: the whole ring is requested before the dependency wait
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_wp) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tm_dpr) : "memory");
    for (int i = 0; i < AB_NA; ++i) {
      mbar_expect_tx(a_full + 8 * i, AB_A_STAGE);
      tma_load_4d(sbase + AB_A_OFF + i * AB_A_STAGE, &p.tm_wp, a_full + 8 * i, 0, 128 * i, m_base / 64, 0);
    }
  }
  pdl_wait();
  pdl_trigger();
  cluster_barrier_all();

  if (warp == 0) {
    if (lane == 0) {
      for (int i = AB_NA; i < AB_H / 128; ++i) {         // 6 boxes of 128 out-feature rows
        const int s = i % AB_NA;
        mbar_wait(a_empty + 8 * s, ((uint32_t)(i / AB_NA) & 1u) ^ 1u);
        mbar_expect_tx(a_full + 8 * s, AB_A_STAGE);
        tma_load_4d(sbase + AB_A_OFF + s * AB_A_STAGE, &p.tm_wp, a_full + 8 * s, 0, 128 * i, m_base / 64, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait_cluster(cl_s, 0);                        // every CTA has stored its slice of dropout(dy1)
      fence_proxy_async();
      mbar_expect_tx(b_full, AB_B_BYTES);
      tma_load_4d(sbase + AB_B_OFF, &p.tm_dpr, b_full, 0, row0, 0, 0);
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // D = f32, A = bf16 MN-major, B = bf16 K-major, N = 64, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | ((uint32_t)(AB_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait(b_full, 0);
      tc_fence_after();
      for (int i = 0; i < AB_H / 128; ++i) {
        const int s = i % AB_NA;
        mbar_wait(a_full + 8 * s, (uint32_t)(i / AB_NA) & 1u);
        tc_fence_after();
        const uint32_t sa = sbase + AB_A_OFF + s * AB_A_STAGE;
#pragma unroll
        for (int c = 0; c < 2; ++c) {                    // two 64-row k-blocks per box
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            // A: [group (64 in-features)][128 k-rows][128 B]: groups 16 KB apart, 16 k-rows = 2 KB
            const uint64_t ad = make_sdesc(sa + c * 8192 + j * 2048, 16384, 1024);
            const uint64_t bd = make_sdesc(sbase + AB_B_OFF + (2 * i + c) * 8192 + j * 32, 16, 1024);
            umma_bf16(tmem, ad, bd, idesc, (i > 0 || c > 0 || j > 0) ? 1u : 0u);
          }
        }
        umma_commit(a_empty + 8 * s);
      }
      umma_commit(tmem_full);
    }
  } else if (warp >= 4) {
    const int cw = warp - 4, ctid = threadIdx.x - 128;
    const int g = lane >> 2, t4 = lane & 3;
    bf16* DYs = reinterpret_cast<bf16*>(smem + AB_DY_OFF);
    float* stat = reinterpret_cast<float*>(smem + AB_STAT_OFF);
    float* gb = reinterpret_cast<float*>(smem + AB_GB_OFF);
    const int64_t MH = (int64_t)M * AB_H;
    // ------------------------------------------------------------------ A. LN1 backward on [64 tokens] x [96 features]
    {
      const int c = ctid >> 2, part = ctid & 3, f0 = 24 * part;
      const bool tok_ok = c < nrows;
      const int64_t goff = (int64_t)(row0 + c) * AB_H + AB_D * (int)h + f0;
      for (int i = ctid; i < 2 * AB_D; i += 256) gb[i] = 0.0f;
      float gv[24], xh[24];
      float mu = 0.0f, rs = 0.0f;
      if (tok_ok) {
        mu = p.mean1[row0 + c];
        rs = p.rstd1[row0 + c];
      }
#pragma unroll
      for (int i = 0; i < 24; ++i) gv[i] = 0.0f;
      if (tok_ok) {
        for (int k = 0; k < p.nparts; ++k) {
          const float4* pr = reinterpret_cast<const float4*>(p.parts + (int64_t)k * p.part_stride + goff);
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 tq = pr[i];
            gv[4 * i] += tq.x; gv[4 * i + 1] += tq.y; gv[4 * i + 2] += tq.z; gv[4 * i + 3] += tq.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        float r8[8], y8[8];
        uint4 ur = make_uint4(0, 0, 0, 0), uy = make_uint4(0, 0, 0, 0);
        if (tok_ok) {
          if (p.dres) ur = *reinterpret_cast<const uint4*>(p.dres + goff + 8 * i);
          uy = *reinterpret_cast<const uint4*>(p.y1 + goff + 8 * i);
        }
        unpack8(ur, r8);
        unpack8(uy, y8);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          gv[8 * i + e] = bf16_round(gv[8 * i + e] + r8[e]);      // what a dgrad GEMM with a residual epilogue stores
          xh[8 * i + e] = (y8[e] - mu) * rs;
        }
      }
      float s1 = 0.0f, s2 = 0.0f;
      const float* gm = p.g1 + AB_D * (int)h + f0;
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        const float gg = gv[i] * __ldg(gm + i);
        s1 += gg;
        s2 += gg * xh[i];
      }
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
      s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
      if (part == 0) {
        const uint32_t local = smem_u32(stat + ((int)h * AB_N + c) * 2);
#pragma unroll
        for (uint32_t r = 0; r < RFC_HEADS; ++r) st_cluster_f32x2(map_to_cta(local, r), s1, s2);
      }
      // dgamma / dbeta of this CTA's features: sum over the 8 tokens of the warp (lanes 4 apart), then shared atomics
      compute_sync();                                    // gb zeroed, remote stores issued
      cluster_publish(cl_e, ctid);
#pragma unroll
      for (int i = 0; i < 24; ++i) {
        float dg = gv[i] * xh[i], db = gv[i];
        dg += __shfl_xor_sync(0xffffffffu, dg, 4);
        dg += __shfl_xor_sync(0xffffffffu, dg, 8);
        dg += __shfl_xor_sync(0xffffffffu, dg, 16);
        db += __shfl_xor_sync(0xffffffffu, db, 4);
        db += __shfl_xor_sync(0xffffffffu, db, 8);
        db += __shfl_xor_sync(0xffffffffu, db, 16);
        if (lane < 4) {
          atomicAdd(gb + f0 + i, dg);
          atomicAdd(gb + AB_D + f0 + i, db);
        }
      }
      mbar_wait_cluster(cl_e, 0);
      float t1 = 0.0f, t2 = 0.0f;
#pragma unroll
      for (int s = 0; s < RFC_HEADS; ++s) {
        t1 += stat[(s * AB_N + c) * 2];
        t2 += stat[(s * AB_N + c) * 2 + 1];
      }
      t1 *= 1.0f / (float)AB_H;
      t2 *= 1.0f / (float)AB_H;
      const unsigned long long sd = seed_eff(p.seed, p.seed_ctr);
      const uint32_t thr = (uint32_t)(p.p1 * 4294967296.0);
      const float keep = p.p1 > 0.0f ? 1.0f / (1.0f - p.p1) : 1.0f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint4 o, od;
        uint32_t* ow = &o.x;
        uint32_t* dw = &od.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float a[2], dd[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int e = 8 * i + 2 * k + u;
            const float gg = gv[e] * __ldg(gm + e);
            a[u] = tok_ok ? bf16_round(rs * (gg - t1 - xh[e] * t2)) : 0.0f;
            dd[u] = a[u];
            if (p.p1 > 0.0f)
              dd[u] = hash32(sd, (uint64_t)(row0 + c) * AB_H + (uint64_t)(AB_D * (int)h + f0 + e)) >= thr ? a[u] * keep : 0.0f;
          }
          ow[k] = pack2(a[0], a[1]);
          dw[k] = pack2(dd[0], dd[1]);
        }
        *reinterpret_cast<uint4*>(DYs + c * AB_LDN + f0 + 8 * i) = o;
        if (tok_ok) *reinterpret_cast<uint4*>(p.dpr + goff + 8 * i) = od;
      }
      fence_proxy_async();
      compute_sync();                                    // DYs complete, gb complete, dpr stores issued
      cluster_publish(cl_s, ctid);
      for (int i = ctid; i < 2 * AB_D; i += 256) {
        float* dst = (i < AB_D ? p.dg1 : p.db1) + AB_D * (int)h + (i % AB_D);
        atomicAdd(dst, gb[i]);
      }
    }
    // ------------------------------------------------------------------ D. dO of this head: TMEM -> mma.sync tiles
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    // the ring and the B region are free now: Wkqv and the per-sample tiles live there
    bf16* Ws = reinterpret_cast<bf16*>(smem);                                    // [288][104]
    auto sample_base = [&](int s) { return smem + AB_WS_BYTES + s * AB_SAMPLE_STRIDE; };
    {
      const int cpr = AB_D / 8, total = 3 * AB_D * cpr;
      for (int idx = ctid; idx < total; idx += 256) {
        const int row = idx / cpr, ch = idx - row * cpr;
        cp_async16(Ws + row * AB_LDN + ch * 8, p.wkqv + row * AB_D + ch * 8);
      }
    }
    {
      uint4* z = reinterpret_cast<uint4*>(smem + AB_WS_BYTES);
      for (int i = ctid; i < 2 * AB_SAMPLE_STRIDE / 16; i += 256) z[i] = make_uint4(0, 0, 0, 0);
    }
    compute_sync();
    {
      // thread (q, lane) holds in-feature 32q + lane of the accumulator; the head's features are lanes lane_off ...
      const int q = cw & 3, chh = cw >> 2;
      const int f = 32 * q + lane - lane_off;
      const uint32_t tmem_lane = tmem + ((uint32_t)(32 * q) << 16);
#pragma unroll 1
      for (int cb = 0; cb < 32; cb += 8) {
        uint32_t r[8];
        __syncwarp();
        tmem_ld8(tmem_lane + chh * 32 + cb, r);
        tmem_ld_wait();
        if (f >= 0 && f < AB_D) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = chh * 32 + cb + k;
            if (c < nrows) {
              const int s = c >= T ? 1 : 0, tq = c - s * T;
              bf16* dOs = reinterpret_cast<bf16*>(sample_base(s)) + AB_TP * AB_LDN;                    // after Vs
              bf16* dOt = dOs + AB_TP * AB_LDN + AB_D * AB_LDT;                                         // after Kt
              const bf16 v = __float2bfloat16_rn(__uint_as_float(r[k]));
              dOs[tq * AB_LDN + f] = v;
              dOt[f * AB_LDT + tq] = v;
            }
          }
        }
      }
    }
    // K, Q, V of the group's samples (saved by the forward pass)
    for (int s = 0; s < nsamp; ++s) {
      bf16* Vs = reinterpret_cast<bf16*>(sample_base(s));
      bf16* dOs = Vs + AB_TP * AB_LDN;
      bf16* Kt = dOs + AB_TP * AB_LDN;
      bf16* dOt = Kt + AB_D * AB_LDT;
      bf16* Qt = dOt + AB_D * AB_LDT;
      const bf16* base = p.kqv + ((int64_t)(s0 + s) * T * RFC_HEADS + h) * (3 * AB_D);
      const int64_t rstride = (int64_t)RFC_HEADS * 3 * AB_D;
      ab_stage_tile(base + 2 * AB_D, rstride, T, AB_D, Vs, AB_LDN, nullptr, 0, ctid, 256);
      ab_stage_tile(base, rstride, T, AB_D, nullptr, 0, Kt, AB_LDT, ctid, 256);
      ab_stage_tile(base + AB_D, rstride, T, AB_D, nullptr, 0, Qt, AB_LDT, ctid, 256);
    }
    compute_sync();
    // ------------------------------------------------------------------ E. residual-attention backward (two strips per sample)
    const int s_me = cw >> 1, r0 = (cw & 1) * 16;
    const bool active = s_me < nsamp && r0 < ((T + 15) & ~15);
    bf16* Vs = reinterpret_cast<bf16*>(sample_base(s_me < 2 ? s_me : 0));
    bf16* dOs = Vs + AB_TP * AB_LDN;
    bf16* Kt = dOs + AB_TP * AB_LDN;
    bf16* dOt = Kt + AB_D * AB_LDT;
    bf16* Qt = dOt + AB_D * AB_LDT;
    bf16* Pt = Qt + AB_D * AB_LDT;
    bf16* dSt = Pt + AB_TP * AB_LDT;
    bf16* Gs = dSt + AB_TP * AB_LDT;
    const int b = s0 + s_me;
    const int64_t sb = ((int64_t)b * RFC_HEADS + h) * T * T;
    const float inv_sqrt_d = 1.0f / sqrtf((float)AB_D);
    constexpr int NT = AB_TP / 8, KS = AB_TP / 16, ND = AB_D / 8;
    bf16* dbase = p.dkqv + ((int64_t)b * T * RFC_HEADS + h) * (3 * AB_D);
    const int64_t drow = (int64_t)RFC_HEADS * 3 * AB_D;
    if (active) {
      const int iA = r0 + g, iB = r0 + g + 8;
      float pr[NT][4];
      float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (e < 2) ? iA : iB;
          const int j = nt * 8 + 2 * t4 + (e & 1);
          float v = -INFINITY;
          if (i < T && j < T) v = __ldg(p.scores + sb + (int64_t)i * T + j);
          pr[nt][e] = v;
          if (e < 2) mxA = fmaxf(mxA, v); else mxB = fmaxf(mxB, v);
        }
      }
      mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1));
      mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
      mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1));
      mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
      float sumA = 0.0f, sumB = 0.0f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float m = (e < 2) ? mxA : mxB;
          const float qv = (pr[nt][e] == -INFINITY) ? 0.0f : expf(pr[nt][e] - m);
          pr[nt][e] = qv;
          if (e < 2) sumA += qv; else sumB += qv;
        }
      }
      sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
      sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
      const float invA = sumA > 0.0f ? 1.0f / sumA : 0.0f, invB = sumB > 0.0f ? 1.0f / sumB : 0.0f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        pr[nt][0] *= invA; pr[nt][1] *= invA; pr[nt][2] *= invB; pr[nt][3] *= invB;
      }
      // dP = dO V^T
      float dp[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.0f;
#pragma unroll
      for (int kd = 0; kd < AB_D / 16; ++kd) {
        uint32_t a[4];
        load_a(a, dOs, AB_LDN, r0, kd * 16, g, t4);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          uint32_t bb[2];
          load_b(bb, Vs, AB_LDN, nt * 8, kd * 16, g, t4);
          mma_bf16_16816(dp[nt], a, bb);
        }
      }
      float rdA = 0.0f, rdB = 0.0f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (e < 2) ? iA : iB;
          const int j = nt * 8 + 2 * t4 + (e & 1);
          Pt[j * AB_LDT + i] = __float2bfloat16_rn(pr[nt][e]);
          const float contrib = pr[nt][e] * dp[nt][e];
          if (e < 2) rdA += contrib; else rdB += contrib;
        }
      }
      rdA += __shfl_xor_sync(0xffffffffu, rdA, 1);
      rdA += __shfl_xor_sync(0xffffffffu, rdA, 2);
      rdB += __shfl_xor_sync(0xffffffffu, rdB, 1);
      rdB += __shfl_xor_sync(0xffffffffu, rdB, 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int i = (e < 2) ? iA : iB;
          const int j = nt * 8 + 2 * t4 + (e & 1);
          float gsc = pr[nt][e] * (dp[nt][e] - ((e < 2) ? rdA : rdB));
          if (i < T && j < T) {
            if (p.ds_in) gsc += __ldg(p.ds_in + sb + (int64_t)i * T + j);
            if (p.dprev) p.dprev[sb + (int64_t)i * T + j] = gsc;
          } else {
            gsc = 0.0f;
          }
          dp[nt][e] = gsc;
          dSt[j * AB_LDT + i] = __float2bfloat16_rn(gsc);
        }
      }
      // dQ = dS K / sqrt(d)
      float dq[ND][4];
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.0f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t a[4];
        a[0] = pack2(dp[2 * ks][0], dp[2 * ks][1]);
        a[1] = pack2(dp[2 * ks][2], dp[2 * ks][3]);
        a[2] = pack2(dp[2 * ks + 1][0], dp[2 * ks + 1][1]);
        a[3] = pack2(dp[2 * ks + 1][2], dp[2 * ks + 1][3]);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          uint32_t bb[2];
          load_b(bb, Kt, AB_LDT, nd * 8, ks * 16, g, t4);
          mma_bf16_16816(dq[nd], a, bb);
        }
      }
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        const int sc = nd * 8 + 2 * t4;
        const uint32_t qa = pack2(dq[nd][0] * inv_sqrt_d, dq[nd][1] * inv_sqrt_d), qb = pack2(dq[nd][2] * inv_sqrt_d, dq[nd][3] * inv_sqrt_d);
        if (iA < T) *reinterpret_cast<uint32_t*>(dbase + AB_D + (int64_t)iA * drow + sc) = qa;
        if (iB < T) *reinterpret_cast<uint32_t*>(dbase + AB_D + (int64_t)iB * drow + sc) = qb;
        *reinterpret_cast<uint32_t*>(Gs + iA * AB_LDG + AB_D + sc) = qa;       // rows >= T hold zeros (dS is zero there)
        *reinterpret_cast<uint32_t*>(Gs + iB * AB_LDG + AB_D + sc) = qb;
      }
    }
    compute_sync();
    if (active) {
      // this warp now owns KEY rows [r0, r0 + 16): dV = P^T dO, dK = dS^T Q / sqrt(d)
      const int jA = r0 + g, jB = r0 + g + 8;
      float dv[ND][4], dk[ND][4];
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.0f;
        dk[nd][0] = dk[nd][1] = dk[nd][2] = dk[nd][3] = 0.0f;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t ap[4], as[4];
        load_a(ap, Pt, AB_LDT, r0, ks * 16, g, t4);
        load_a(as, dSt, AB_LDT, r0, ks * 16, g, t4);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          uint32_t bb[2];
          load_b(bb, dOt, AB_LDT, nd * 8, ks * 16, g, t4);
          mma_bf16_16816(dv[nd], ap, bb);
          load_b(bb, Qt, AB_LDT, nd * 8, ks * 16, g, t4);
          mma_bf16_16816(dk[nd], as, bb);
        }
      }
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        const int sc = nd * 8 + 2 * t4;
        const uint32_t va = pack2(dv[nd][0], dv[nd][1]), vb = pack2(dv[nd][2], dv[nd][3]);
        const uint32_t ka = pack2(dk[nd][0] * inv_sqrt_d, dk[nd][1] * inv_sqrt_d), kb = pack2(dk[nd][2] * inv_sqrt_d, dk[nd][3] * inv_sqrt_d);
        if (jA < T) {
          *reinterpret_cast<uint32_t*>(dbase + 2 * AB_D + (int64_t)jA * drow + sc) = va;
          *reinterpret_cast<uint32_t*>(dbase + (int64_t)jA * drow + sc) = ka;
        }
        if (jB < T) {
          *reinterpret_cast<uint32_t*>(dbase + 2 * AB_D + (int64_t)jB * drow + sc) = vb;
          *reinterpret_cast<uint32_t*>(dbase + (int64_t)jB * drow + sc) = kb;
        }
        *reinterpret_cast<uint32_t*>(Gs + jA * AB_LDG + 2 * AB_D + sc) = va;   // padded key rows: zeros
        *reinterpret_cast<uint32_t*>(Gs + jB * AB_LDG + 2 * AB_D + sc) = vb;
        *reinterpret_cast<uint32_t*>(Gs + jA * AB_LDG + sc) = ka;
        *reinterpret_cast<uint32_t*>(Gs + jB * AB_LDG + sc) = kb;
      }
    }
    // ------------------------------------------------------------------ kqv dgrad: dx = dkqv . Wkqv + dy1, all 8 warps
    cp_async_wait_all();
    compute_sync();
    {
      const int ntile_n = AB_D / 8, strips = AB_TP / 16, kk_n = 3 * AB_D / 16;
      const int npairs = nsamp * strips * ntile_n;
      for (int pi = cw; pi < npairs; pi += 8) {
        const int s = pi / (strips * ntile_n), rem = pi - s * strips * ntile_n;
        const int sidx = rem / ntile_n, nt = rem - sidx * ntile_n;
        const bf16* Gss = reinterpret_cast<const bf16*>(sample_base(s)) + 2 * AB_TP * AB_LDN + 3 * AB_D * AB_LDT + 2 * AB_TP * AB_LDT;
        float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 6
        for (int kk = 0; kk < kk_n; ++kk) {
          uint32_t a[4], bb[2];
          load_a(a, Gss, AB_LDG, sidx * 16, kk * 16, g, t4);
          load_b_trans(bb, Ws, AB_LDN, kk * 16, nt * 8, lane);
          mma_bf16_16816(c, a, bb);
        }
        const int n = nt * 8 + 2 * t4;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int tq = sidx * 16 + g + 8 * half;
          if (tq < T) {
            const int cidx = s * T + tq;                 // column of the group
            const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(DYs + cidx * AB_LDN + n);
            *reinterpret_cast<uint32_t*>(p.dxin + (int64_t)(row0 + cidx) * AB_H + AB_D * (int)h + n) =
                pack2(c[2 * half] + __bfloat162float(rr.x), c[2 * half + 1] + __bfloat162float(rr.y));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_barrier_all();
  if (warp == 3) tmem_dealloc(tmem, 64);
}

static int ab_samples_per_cluster(int T) { return AB_N / T < 2 ? AB_N / T : 2; }

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_rf_attn_block_bwd_supported(int B, int T, int hidden, int heads) {
  if (hidden != AB_H || heads != RFC_HEADS || T < 1 || T > AB_TP || B < 1) return 0;
  const int spc = ab_samples_per_cluster(T);
  return (B + spc - 1) / spc <= 14 ? 1 : 0;
}

int mmvqa_rf_attn_block_bwd(const mmvqa_rf_attn_block_bwd_args* a, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(a != nullptr, "rf_attn_block_bwd: null args");
  MMVQA_REQUIRE(mmvqa_rf_attn_block_bwd_supported(a->B, a->T, a->hidden, a->heads),
                "rf_attn_block_bwd: unsupported shape (needs hidden 768, 8 heads, T <= 32, <= 14 sample groups)");
  MMVQA_REQUIRE(a->dy_parts && a->nparts >= 1 && a->y1 && a->mean1 && a->rstd1 && a->ln1_w && a->wproj && a->wkqv && a->kqv &&
                    a->scores && a->dpr && a->dkqv && a->dxin && a->dln1_w && a->dln1_b, "rf_attn_block_bwd: null buffer");
  MMVQA_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "rf_attn_block_bwd: dropout must be in [0,1)");
  const int M = a->B * a->T;
  MMVQA_REQUIRE(a->part_stride >= (int64_t)M * AB_H && a->part_stride % 4 == 0, "rf_attn_block_bwd: bad slab stride");
  const uintptr_t al = reinterpret_cast<uintptr_t>(a->dy_parts) | reinterpret_cast<uintptr_t>(a->y1) | reinterpret_cast<uintptr_t>(a->dy_res) |
                       reinterpret_cast<uintptr_t>(a->dpr) | reinterpret_cast<uintptr_t>(a->wkqv) | reinterpret_cast<uintptr_t>(a->kqv) |
                       reinterpret_cast<uintptr_t>(a->dkqv) | reinterpret_cast<uintptr_t>(a->dxin) | reinterpret_cast<uintptr_t>(a->wproj);
  MMVQA_REQUIRE((al & 15) == 0, "rf_attn_block_bwd: buffers must be 16-byte aligned");
  static RfAttnBwdParams P;
  int rc;
  // Wp [out = 768 rows, in = 768 contiguous]: the in-features are the UMMA M dimension (MN-major A), two 64-wide groups
  if ((rc = tc_make_map_chunked(&P.tm_wp, a->wproj, AB_H, AB_H, AB_H, 1, 0, 128, 2, "proj.weight"))) return rc;
  if ((rc = tc_make_map_chunked(&P.tm_dpr, a->dpr, AB_H, M, AB_H, 1, 0, AB_N, AB_H / 64, "dropout(dy1)"))) return rc;
  P.parts = a->dy_parts; P.nparts = a->nparts; P.part_stride = a->part_stride;
  P.dres = reinterpret_cast<const bf16*>(a->dy_res);
  P.y1 = reinterpret_cast<const bf16*>(a->y1); P.mean1 = a->mean1; P.rstd1 = a->rstd1; P.g1 = a->ln1_w;
  P.wkqv = reinterpret_cast<const bf16*>(a->wkqv); P.kqv = reinterpret_cast<const bf16*>(a->kqv);
  P.scores = a->scores; P.ds_in = a->dscores_in;
  P.dpr = reinterpret_cast<bf16*>(a->dpr); P.dkqv = reinterpret_cast<bf16*>(a->dkqv); P.dprev = a->dprev;
  P.dxin = reinterpret_cast<bf16*>(a->dxin); P.dg1 = a->dln1_w; P.db1 = a->dln1_b;
  P.B = a->B; P.T = a->T; P.M = M; P.spc = ab_samples_per_cluster(a->T);
  P.p1 = a->dropout_p; P.seed = a->dropout_seed; P.seed_ctr = g_seed_ctr;
  const int groups = (a->B + P.spc - 1) / P.spc;
  static bool attr_set = false;
  if (!attr_set) {
    MMVQA_CUDA(cudaFuncSetAttribute(rf_attn_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(groups * RFC_HEADS);
  cfg.blockDim = dim3(AB_THREADS);
  cfg.dynamicSmemBytes = AB_SMEM;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  MMVQA_CUDA(cudaLaunchKernelEx(&cfg, rf_attn_block_bwd_kernel, P));
  MMVQA_LAUNCHED("rf_attn_block_bwd");
  return MMVQA_OK;
}

}  // extern "C"
