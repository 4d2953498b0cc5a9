// attention_tc.cuh -- bf16 tensor-core variant of the fused short-sequence attention kernels (T <= 128,
// d % 16 == 0, d <= 128).  One CTA per (batch, head); one warp per 16-query-row strip.
//   forward : S = Q K^T (mma.sync m16n8k16, fp32 accumulators stay in registers) -> scale, + prev, mask in
//             registers -> scores written once (fp32) -> warp-shuffle softmax on the accumulator fragments -> the P
//             fragments are re-packed in registers as the A operand of O = P V (V staged transposed) -> bf16 out.
//   backward: per strip P (recomputed from the stored scores), dP = dO V^T, dS = P (dP - rowdot) + dS_next (the
//             running RealFormer score gradient, written once as dprev), dQ = dS K from registers; P and dS strips
//             are parked transposed in shared memory, then the warps re-partition over key strips for
//             dV = P^T dO and dK = dS^T Q.
// The T x T tile never exists in shared or global memory except for the tensors the reference itself materialises.
#pragma once
#include "common.cuh"

namespace mmvqa {

struct AttnLayout {
  int64_t row_stride;   // elements between consecutive tokens of the same head
  int64_t head_stride;  // elements between heads of the same token
  int64_t tok_batch;    // elements between batches (= T * row_stride)
  int q_off, k_off, v_off;
};

using bf16 = __nv_bfloat16;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x (low 16 bits) = lo
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds32(const bf16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// A fragment (16 rows x 16 k) from a row-major [row][k] bf16 tile
__device__ __forceinline__ void load_a(uint32_t (&a)[4], const bf16* X, int ld, int r0, int k0, int g, int t) {
  a[0] = lds32(X + (r0 + g) * ld + k0 + 2 * t);
  a[1] = lds32(X + (r0 + g + 8) * ld + k0 + 2 * t);
  a[2] = lds32(X + (r0 + g) * ld + k0 + 2 * t + 8);
  a[3] = lds32(X + (r0 + g + 8) * ld + k0 + 2 * t + 8);
}
// B fragment (16 k x 8 n) from a [n][k] bf16 tile (k contiguous)
__device__ __forceinline__ void load_b(uint32_t (&b)[2], const bf16* Y, int ld, int n0, int k0, int g, int t) {
  b[0] = lds32(Y + (n0 + g) * ld + k0 + 2 * t);
  b[1] = lds32(Y + (n0 + g) * ld + k0 + 2 * t + 8);
}

// global [Tn, d] bf16 tile (row stride in elements) -> natural [row][d+8] and/or transposed [col][Tp+8] smem copies
__device__ __forceinline__ void stage_tile(const bf16* __restrict__ src, int64_t row_stride, int Tn, int d, bf16* nat,
                                           int ldn, bf16* tr, int ldt) {
  const int cpr = d / 8;
  const int total = Tn * cpr;
  for (int base = threadIdx.x; base < total; base += blockDim.x * 4) {
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * blockDim.x;
      if (idx < total) {
        const int row = idx / cpr, ch = idx - row * cpr;
        v[k] = *reinterpret_cast<const uint4*>(src + row * row_stride + ch * 8);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int idx = base + k * blockDim.x;
      if (idx < total) {
        const int row = idx / cpr, ch = idx - row * cpr;
        if (nat) *reinterpret_cast<uint4*>(nat + row * ldn + ch * 8) = v[k];
        if (tr) {
          const uint32_t w[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const unsigned short bits = (unsigned short)((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
            reinterpret_cast<unsigned short*>(tr)[(ch * 8 + e) * ldt + row] = bits;
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void zero_smem(void* base, int bytes) {
  uint4* p = reinterpret_cast<uint4*>(base);
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) p[i] = z;
}

// B fragment (16 k x 8 n) of m16n8k16 from a ROW-MAJOR [k][n] bf16 tile (n contiguous): ldmatrix .trans hands every
// thread (k = 2t, 2t+1 | n = g) of each 8x8 block -- exactly the col-major B layout, no transposed copy of the tile
__device__ __forceinline__ void load_b_trans(uint32_t (&b)[2], const bf16* W, int ld, int k0, int n0, int lane) {
  const bf16* row = W + (k0 + (lane & 15)) * ld + n0;        // lanes 0-7: rows k0..k0+7, lanes 8-15: rows k0+8..k0+15
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(addr));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
// fire-and-forget copy of a global [rows, cols] bf16 tile (cols % 8 == 0) into a [rows][ld] smem tile: every thread
// issues all of its 16-byte cp.async requests back to back, ONE memory latency for the whole tile
__device__ __forceinline__ void stage_async(const bf16* __restrict__ src, int64_t row_stride, int rows, int cols, bf16* dst,
                                            int ld) {
  const int cpr = cols / 8, total = rows * cpr;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int row = idx / cpr, ch = idx - row * cpr;
    cp_async16(dst + row * ld + ch * 8, src + row * row_stride + ch * 8);
  }
}

// ---------------------------------------------------------------------------------------------------
// forward.  TP = register-array bound on the padded sequence (32 / 64 / 96 / 128)
// ---------------------------------------------------------------------------------------------------
// FUSED (RealFormer): the shared-weight kqv projection (realformer.py:13,33) runs inside this kernel --
// kqv_bh = x[b, :, h d:(h+1) d] . Wkqv^T  ([T, d] x [d, 3d], mma.sync), written once to global memory for the backward
// pass and straight into the Q / K / V^T shared-memory tiles; `qkv` is then the OUTPUT buffer and the separate
// [M*heads, 3d] x [3d, d] GEMM launch disappears.  Wkqv (a weight) is requested before griddepcontrol.wait.
template <bool RF, int TP, bool FUSED = false>
__global__ void __launch_bounds__(256) attn_tc_fwd_kernel(const bf16* __restrict__ qkv, AttnLayout L,
                                                          const float* __restrict__ prev, const float* __restrict__ mask,
                                                          bf16* __restrict__ out, float* __restrict__ scores,
                                                          bf16* __restrict__ probs, int Tn, int heads, int d, float drop_p,
                                                          unsigned long long seed, const unsigned long long* seed_ctr,
                                                          const bf16* __restrict__ xin = nullptr,
                                                          const bf16* __restrict__ wkqv = nullptr) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int NT = TP / 8, KS = TP / 16;
  const int Tp = (Tn + 15) & ~15;            // rows / keys padded to the mma tile
  const int ldn = d + 8, ldt = Tp + 8;
  bf16* Qs = reinterpret_cast<bf16*>(smem_attn);     // [Tp][d+8]
  bf16* Ks = Qs + Tp * ldn;                          // [Tp][d+8]
  bf16* Vt = Ks + Tp * ldn;                          // [d][Tp+8]   V transposed: B operand of P V
  const int smem_bytes = (2 * Tp * ldn + d * ldt) * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const bf16* base = qkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  if (FUSED) {
    bf16* Xs = Vt + d * ldt;                         // [Tp][d+8]   this head's slice of the layer input
    bf16* Ws = Xs + Tp * ldn;                        // [3d][d+8]   kqv weight, rows k | q | v
    zero_smem(smem_attn, smem_bytes + Tp * ldn * 2);
    __syncthreads();
    stage_async(wkqv, d, 3 * d, d, Ws, ldn);         // weights do not depend on the previous kernel
    pdl_wait();
    pdl_trigger();
    stage_async(xin + (int64_t)b * Tn * heads * d + (int64_t)h * d, (int64_t)heads * d, Tn, d, Xs, ldn);
    cp_async_wait_all();
    __syncthreads();
    const int nwarps = blockDim.x >> 5, ntile_n = 3 * d / 8, npairs = (Tp / 16) * ntile_n;
    bf16* kqv_out = const_cast<bf16*>(base);
    for (int pi = warp; pi < npairs; pi += nwarps) {
      const int sidx = pi / ntile_n, nt = pi - sidx * ntile_n;
      float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int kd = 0; kd < d / 16; ++kd) {
        uint32_t a[4], bb[2];
        load_a(a, Xs, ldn, sidx * 16, kd * 16, g, t);
        load_b(bb, Ws, ldn, nt * 8, kd * 16, g, t);
        mma_bf16_16816(c, a, bb);
      }
      const int rA = sidx * 16 + g, rB = rA + 8;
      const int n = nt * 8 + 2 * t;                  // even: the pair (n, n+1) never straddles k | q | v (d % 8 == 0)
      const int sec = n / d, cc = n - sec * d;
      const uint32_t vA = pack2(c[0], c[1]), vB = pack2(c[2], c[3]);
      if (sec == 2) {                                // V is consumed transposed
        const __nv_bfloat162 pa = *reinterpret_cast<const __nv_bfloat162*>(&vA), pb = *reinterpret_cast<const __nv_bfloat162*>(&vB);
        Vt[cc * ldt + rA] = pa.x; Vt[(cc + 1) * ldt + rA] = pa.y;
        Vt[cc * ldt + rB] = pb.x; Vt[(cc + 1) * ldt + rB] = pb.y;
      } else {
        bf16* dst = (L.k_off == sec * d) ? Ks : Qs;  // section order in the packed row is given by the layout offsets
        *reinterpret_cast<uint32_t*>(dst + rA * ldn + cc) = vA;
        *reinterpret_cast<uint32_t*>(dst + rB * ldn + cc) = vB;
      }
      if (rA < Tn) *reinterpret_cast<uint32_t*>(kqv_out + (int64_t)rA * L.row_stride + n) = vA;
      if (rB < Tn) *reinterpret_cast<uint32_t*>(kqv_out + (int64_t)rB * L.row_stride + n) = vB;
    }
    __syncthreads();
  } else {
    zero_smem(smem_attn, smem_bytes);
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    stage_tile(base + L.q_off, L.row_stride, Tn, d, Qs, ldn, nullptr, 0);
    stage_tile(base + L.k_off, L.row_stride, Tn, d, Ks, ldn, nullptr, 0);
    stage_tile(base + L.v_off, L.row_stride, Tn, d, nullptr, 0, Vt, ldt);
    __syncthreads();
  }
  const int r0 = warp * 16;
  if (r0 >= Tp) return;
  const int nt_n = Tp / 8, ks_n = Tp / 16, kd_n = d / 16, nd_n = d / 8;
  const int iA = r0 + g, iB = r0 + g + 8;
  const int64_t sbase = ((int64_t)b * heads + h) * Tn * Tn;
  const int H = heads * d;
  // ---- S = Q K^T
  float s[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
  for (int kd = 0; kd < kd_n; ++kd) {
    uint32_t a[4];
    load_a(a, Qs, ldn, r0, kd * 16, g, t);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt < nt_n) {
        uint32_t bb[2];
        load_b(bb, Ks, ldn, nt * 8, kd * 16, g, t);
        mma_bf16_16816(s[nt], a, bb);
      }
    }
  }
  // ---- scale, residual scores, mask, write scores, softmax (rows iA: c0,c1; iB: c2,c3; cols j0, j0+1)
  const float inv_sqrt_d = 1.0f / sqrtf((float)d);
  float qoffA = 0.0f, qoffB = 0.0f;
  if (RF && mask) {
    if (iA < Tn) qoffA = -10000.0f * (1.0f - __ldg(mask + b * Tn + iA));
    if (iB < Tn) qoffB = -10000.0f * (1.0f - __ldg(mask + b * Tn + iB));
  }
  float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt < nt_n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? iA : iB;
        const int j = nt * 8 + 2 * t + (e & 1);
        float v = -INFINITY;
        if (i < Tn && j < Tn) {
          v = s[nt][e] * inv_sqrt_d;
          if (RF) {
            if (prev) v += __ldg(prev + sbase + (int64_t)i * Tn + j);
            v += (e < 2) ? qoffA : qoffB;
            scores[sbase + (int64_t)i * Tn + j] = v;
          } else if (mask) {
            v -= 10000.0f * (1.0f - __ldg(mask + b * Tn + j));
          }
        }
        s[nt][e] = v;
        if (e < 2) mxA = fmaxf(mxA, v); else mxB = fmaxf(mxB, v);
      }
    }
  }
  mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 1));
  mxA = fmaxf(mxA, __shfl_xor_sync(0xffffffffu, mxA, 2));
  mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 1));
  mxB = fmaxf(mxB, __shfl_xor_sync(0xffffffffu, mxB, 2));
  float sumA = 0.0f, sumB = 0.0f;
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt < nt_n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float m = (e < 2) ? mxA : mxB;
        const float p = (s[nt][e] == -INFINITY) ? 0.0f : expf(s[nt][e] - m);
        s[nt][e] = p;
        if (e < 2) sumA += p; else sumB += p;
      }
    }
  }
  sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
  sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
  sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
  sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
  const float invA = sumA > 0.0f ? 1.0f / sumA : 0.0f, invB = sumB > 0.0f ? 1.0f / sumB : 0.0f;
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  if (!RF && drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    if (nt < nt_n) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? iA : iB;
        const int j = nt * 8 + 2 * t + (e & 1);
        float p = s[nt][e] * ((e < 2) ? invA : invB);
        if (!RF && i < Tn && j < Tn) {
          probs[sbase + (int64_t)i * Tn + j] = __float2bfloat16_rn(p);
          if (drop_p > 0.0f) p = hash32(seed, (uint64_t)(sbase + (int64_t)i * Tn + j)) >= thr ? p * inv_keep : 0.0f;
        }
        s[nt][e] = p;
      }
    }
  }
  // ---- O = P V : the accumulator fragments of P are the A fragments of the second GEMM
  constexpr int ND = 16;                       // d <= 128
  float o[ND][4];
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.0f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    if (ks < ks_n) {
      uint32_t a[4];
      a[0] = pack2(s[2 * ks][0], s[2 * ks][1]);
      a[1] = pack2(s[2 * ks][2], s[2 * ks][3]);
      a[2] = pack2(s[2 * ks + 1][0], s[2 * ks + 1][1]);
      a[3] = pack2(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        if (nd < nd_n) {
          uint32_t bb[2];
          load_b(bb, Vt, ldt, nd * 8, ks * 16, g, t);
          mma_bf16_16816(o[nd], a, bb);
        }
      }
    }
  }
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    if (nd < nd_n) {
      const int sc = nd * 8 + 2 * t;
      if (iA < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * Tn + iA) * H + h * d + sc) = pack2(o[nd][0], o[nd][1]);
      if (iB < Tn) *reinterpret_cast<uint32_t*>(out + ((int64_t)b * Tn + iB) * H + h * d + sc) = pack2(o[nd][2], o[nd][3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------
// FUSED (RealFormer): the input gradient of the shared-weight kqv projection is produced here as well --
// dx[b, :, h d:(h+1) d] = dkqv_bh . Wkqv + dres  ([T, 3d] x [3d, d], mma.sync; dres = the residual gradient around the
// attention block) -- so the [M*heads, 3d] x [3d, d] dgrad GEMM launch disappears from the backward chain.  dkqv is
// still written to global memory: the weight-gradient GEMM on the side branch reads it.
template <bool RF, int TP, bool FUSED = false>
__global__ void __launch_bounds__(256) attn_tc_bwd_kernel(const bf16* __restrict__ qkv, AttnLayout L,
                                                          const float* __restrict__ scores, const bf16* __restrict__ probs,
                                                          const bf16* __restrict__ dout, const float* __restrict__ dscores_in,
                                                          bf16* __restrict__ dqkv, float* __restrict__ dprev, int Tn, int heads,
                                                          int d, float drop_p, unsigned long long seed,
                                                          const unsigned long long* seed_ctr,
                                                          const bf16* __restrict__ wkqv = nullptr,
                                                          const bf16* __restrict__ dres = nullptr, bf16* __restrict__ dx = nullptr) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  constexpr int NT = TP / 8, KS = TP / 16, ND = 16;
  const int Tp = (Tn + 15) & ~15;
  const int ldn = d + 8, ldt = Tp + 8;
  bf16* Vs = reinterpret_cast<bf16*>(smem_attn);     // [Tp][d+8]   B of dP = dO V^T
  bf16* dOs = Vs + Tp * ldn;                         // [Tp][d+8]   A of dP
  bf16* Kt = dOs + Tp * ldn;                         // [d][Tp+8]   B of dQ = dS K
  bf16* dOt = Kt + d * ldt;                          // [d][Tp+8]   B of dV = P^T dO
  bf16* Qt = dOt + d * ldt;                          // [d][Tp+8]   B of dK = dS^T Q
  bf16* Pt = Qt + d * ldt;                           // [Tp][Tp+8]  P^T  (A of dV)
  bf16* dSt = Pt + Tp * ldt;                         // [Tp][Tp+8]  dS^T (A of dK)
  const int ldg = 3 * d + 8;
  bf16* Gs = dSt + Tp * ldt;                         // FUSED: [Tp][3d+8]  dkqv of this (batch, head), A of dx
  bf16* Ws = Gs + Tp * ldg;                          // FUSED: [3d][d+8]   kqv weight, row-major [k][n]
  const int smem_bytes = (2 * Tp * ldn + 3 * d * ldt + 2 * Tp * ldt + (FUSED ? Tp * ldg : 0)) * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int H = heads * d;
  const bf16* base = qkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  bf16* dbase = dqkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  const int64_t sbase = ((int64_t)b * heads + h) * Tn * Tn;
  zero_smem(smem_attn, smem_bytes);
  __syncthreads();
  if (FUSED) stage_async(wkqv, d, 3 * d, d, Ws, ldn);   // weights: in flight before the dependency wait, needed last
  pdl_wait();
  pdl_trigger();
  if (!RF && drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  stage_tile(base + L.v_off, L.row_stride, Tn, d, Vs, ldn, nullptr, 0);
  stage_tile(dout + (int64_t)b * Tn * H + h * d, (int64_t)H, Tn, d, dOs, ldn, dOt, ldt);
  stage_tile(base + L.k_off, L.row_stride, Tn, d, nullptr, 0, Kt, ldt);
  stage_tile(base + L.q_off, L.row_stride, Tn, d, nullptr, 0, Qt, ldt);
  __syncthreads();
  const int r0 = warp * 16;
  const int nt_n = Tp / 8, ks_n = Tp / 16, kd_n = d / 16, nd_n = d / 8;
  const float inv_sqrt_d = 1.0f / sqrtf((float)d);
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  if (r0 < Tp) {
    const int iA = r0 + g, iB = r0 + g + 8;
    // ---- P strip
    float p[NT][4];
    float mxA = -INFINITY, mxB = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? iA : iB;
        const int j = nt * 8 + 2 * t + (e & 1);
        float v = RF ? -INFINITY : 0.0f;
        if (nt < nt_n && i < Tn && j < Tn) {
          if (RF) v = __ldg(scores + sbase + (int64_t)i * Tn + j);
          else v = __bfloat162float(probs[sbase + (int64_t)i * Tn + j]);
        }
        p[nt][e] = v;
        if (RF) { if (e < 2) mxA = fmaxf(mxA, v); else mxB = fmaxf(mxB, v); }
      }
    }
    if (RF) {
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
          const float q = (p[nt][e] == -INFINITY) ? 0.0f : expf(p[nt][e] - m);
          p[nt][e] = q;
          if (e < 2) sumA += q; else sumB += q;
        }
      }
      sumA += __shfl_xor_sync(0xffffffffu, sumA, 1);
      sumA += __shfl_xor_sync(0xffffffffu, sumA, 2);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, 1);
      sumB += __shfl_xor_sync(0xffffffffu, sumB, 2);
      const float invA = sumA > 0.0f ? 1.0f / sumA : 0.0f, invB = sumB > 0.0f ? 1.0f / sumB : 0.0f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        p[nt][0] *= invA; p[nt][1] *= invA; p[nt][2] *= invB; p[nt][3] *= invB;
      }
    }
    // ---- dP = dO V^T
    float dp[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.0f;
    for (int kd = 0; kd < kd_n; ++kd) {
      uint32_t a[4];
      load_a(a, dOs, ldn, r0, kd * 16, g, t);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        if (nt < nt_n) {
          uint32_t bb[2];
          load_b(bb, Vs, ldn, nt * 8, kd * 16, g, t);
          mma_bf16_16816(dp[nt], a, bb);
        }
      }
    }
    // ---- dS = P (dP - rowdot) + dS_next ; park P^T (post-dropout for MHSA) and dS^T
    float rdA = 0.0f, rdB = 0.0f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = (e < 2) ? iA : iB;
        const int j = nt * 8 + 2 * t + (e & 1);
        float pd = p[nt][e];
        if (!RF && drop_p > 0.0f && nt < nt_n && i < Tn && j < Tn) {
          const bool keep = hash32(seed, (uint64_t)(sbase + (int64_t)i * Tn + j)) >= thr;
          dp[nt][e] = keep ? dp[nt][e] * inv_keep : 0.0f;
          pd = keep ? pd * inv_keep : 0.0f;
        }
        if (nt < nt_n) Pt[j * ldt + i] = __float2bfloat16_rn(pd);
        const float contrib = p[nt][e] * dp[nt][e];
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
        const int j = nt * 8 + 2 * t + (e & 1);
        float gsc = p[nt][e] * (dp[nt][e] - ((e < 2) ? rdA : rdB));
        if (nt < nt_n && i < Tn && j < Tn) {
          if (RF) {
            if (dscores_in) gsc += __ldg(dscores_in + sbase + (int64_t)i * Tn + j);
            if (dprev) dprev[sbase + (int64_t)i * Tn + j] = gsc;
          }
        } else {
          gsc = 0.0f;
        }
        dp[nt][e] = gsc;                                   // dp now holds dS
        if (nt < nt_n) dSt[j * ldt + i] = __float2bfloat16_rn(gsc);
      }
    }
    // ---- dQ = dS K / sqrt(d)  (A from registers)
    float dq[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.0f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      if (ks < ks_n) {
        uint32_t a[4];
        a[0] = pack2(dp[2 * ks][0], dp[2 * ks][1]);
        a[1] = pack2(dp[2 * ks][2], dp[2 * ks][3]);
        a[2] = pack2(dp[2 * ks + 1][0], dp[2 * ks + 1][1]);
        a[3] = pack2(dp[2 * ks + 1][2], dp[2 * ks + 1][3]);
#pragma unroll
        for (int nd = 0; nd < ND; ++nd) {
          if (nd < nd_n) {
            uint32_t bb[2];
            load_b(bb, Kt, ldt, nd * 8, ks * 16, g, t);
            mma_bf16_16816(dq[nd], a, bb);
          }
        }
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      if (nd < nd_n) {
        const int sc = nd * 8 + 2 * t;
        const uint32_t qa = pack2(dq[nd][0] * inv_sqrt_d, dq[nd][1] * inv_sqrt_d), qb = pack2(dq[nd][2] * inv_sqrt_d, dq[nd][3] * inv_sqrt_d);
        if (iA < Tn) *reinterpret_cast<uint32_t*>(dbase + L.q_off + (int64_t)iA * L.row_stride + sc) = qa;
        if (iB < Tn) *reinterpret_cast<uint32_t*>(dbase + L.q_off + (int64_t)iB * L.row_stride + sc) = qb;
        if (FUSED) {   // rows >= Tn hold zeros (dS is zero there)
          *reinterpret_cast<uint32_t*>(Gs + iA * ldg + L.q_off + sc) = qa;
          *reinterpret_cast<uint32_t*>(Gs + iB * ldg + L.q_off + sc) = qb;
        }
      }
    }
  }
  __syncthreads();
  if (!FUSED && r0 >= Tp) return;
  // ---- phase 2: this warp now owns KEY rows j in [r0, r0+16): dV = P^T dO, dK = dS^T Q / sqrt(d)
  if (r0 < Tp) {
    const int jA = r0 + g, jB = r0 + g + 8;
    float dv[ND][4], dk[ND][4];
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      dv[nd][0] = dv[nd][1] = dv[nd][2] = dv[nd][3] = 0.0f;
      dk[nd][0] = dk[nd][1] = dk[nd][2] = dk[nd][3] = 0.0f;
    }
    for (int ks = 0; ks < ks_n; ++ks) {
      uint32_t ap[4], as[4];
      load_a(ap, Pt, ldt, r0, ks * 16, g, t);
      load_a(as, dSt, ldt, r0, ks * 16, g, t);
#pragma unroll
      for (int nd = 0; nd < ND; ++nd) {
        if (nd < nd_n) {
          uint32_t bb[2];
          load_b(bb, dOt, ldt, nd * 8, ks * 16, g, t);
          mma_bf16_16816(dv[nd], ap, bb);
          load_b(bb, Qt, ldt, nd * 8, ks * 16, g, t);
          mma_bf16_16816(dk[nd], as, bb);
        }
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      if (nd < nd_n) {
        const int sc = nd * 8 + 2 * t;
        const uint32_t va = pack2(dv[nd][0], dv[nd][1]), vb = pack2(dv[nd][2], dv[nd][3]);
        const uint32_t ka = pack2(dk[nd][0] * inv_sqrt_d, dk[nd][1] * inv_sqrt_d), kb = pack2(dk[nd][2] * inv_sqrt_d, dk[nd][3] * inv_sqrt_d);
        if (jA < Tn) {
          *reinterpret_cast<uint32_t*>(dbase + L.v_off + (int64_t)jA * L.row_stride + sc) = va;
          *reinterpret_cast<uint32_t*>(dbase + L.k_off + (int64_t)jA * L.row_stride + sc) = ka;
        }
        if (jB < Tn) {
          *reinterpret_cast<uint32_t*>(dbase + L.v_off + (int64_t)jB * L.row_stride + sc) = vb;
          *reinterpret_cast<uint32_t*>(dbase + L.k_off + (int64_t)jB * L.row_stride + sc) = kb;
        }
        if (FUSED) {   // padded key rows: P^T / dS^T rows are zero there, so these are zeros
          *reinterpret_cast<uint32_t*>(Gs + jA * ldg + L.v_off + sc) = va;
          *reinterpret_cast<uint32_t*>(Gs + jB * ldg + L.v_off + sc) = vb;
          *reinterpret_cast<uint32_t*>(Gs + jA * ldg + L.k_off + sc) = ka;
          *reinterpret_cast<uint32_t*>(Gs + jB * ldg + L.k_off + sc) = kb;
        }
      }
    }
  }
  if (FUSED) {
    // ---- phase 3: dx = dkqv . Wkqv + dres over all warps; (16-row strip, 8-column tile) pairs round-robin
    cp_async_wait_all();
    __syncthreads();
    const int nwarps = blockDim.x >> 5, ntile_n = d / 8, npairs = (Tp / 16) * ntile_n, kk_n = 3 * d / 16;
    const int64_t xoff = (int64_t)b * Tn * H + (int64_t)h * d;
    for (int pi = warp; pi < npairs; pi += nwarps) {
      const int sidx = pi / ntile_n, nt = pi - sidx * ntile_n;
      float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int kk = 0; kk < kk_n; ++kk) {
        uint32_t a[4], bb[2];
        load_a(a, Gs, ldg, sidx * 16, kk * 16, g, t);
        load_b_trans(bb, Ws, ldn, kk * 16, nt * 8, lane);
        mma_bf16_16816(c, a, bb);
      }
      const int rA = sidx * 16 + g, rB = rA + 8, n = nt * 8 + 2 * t;
      if (rA < Tn) {
        const int64_t o = xoff + (int64_t)rA * H + n;
        float r0_ = 0.0f, r1_ = 0.0f;
        if (dres) { const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(dres + o); r0_ = __bfloat162float(rr.x); r1_ = __bfloat162float(rr.y); }
        *reinterpret_cast<uint32_t*>(dx + o) = pack2(c[0] + r0_, c[1] + r1_);
      }
      if (rB < Tn) {
        const int64_t o = xoff + (int64_t)rB * H + n;
        float r0_ = 0.0f, r1_ = 0.0f;
        if (dres) { const __nv_bfloat162 rr = *reinterpret_cast<const __nv_bfloat162*>(dres + o); r0_ = __bfloat162float(rr.x); r1_ = __bfloat162float(rr.y); }
        *reinterpret_cast<uint32_t*>(dx + o) = pack2(c[2] + r0_, c[3] + r1_);
      }
    }
  }
}

}  // namespace mmvqa
