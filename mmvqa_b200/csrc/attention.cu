// attention.cu -- fused short-sequence attention (T <= 128), one CTA per (batch, head).
//   RF = true : RealFormer residual attention, models/realformer.py:33-44
//               S = q.k/sqrt(d) + prev - 10000*(1 - mask[b,i])   (QUERY-row offset, carried in S)
//               writes S (fp32, the tensor the next layer receives), P = softmax_j S, out = P v
//   RF = false: Transformer MHSA, models/transformer.py:21-27
//               S = q.k/sqrt(d) - 10000*(1 - mask[b,j])          (KEY mask)
//               writes P (the module's self.scores), optional dropout on P, out = P v
// K/Q/V of the (batch, head) are staged once in shared memory as fp32; the T x T score tile never
// leaves the SM except for the tensors the reference itself materialises; softmax is a
// warp-shuffle reduction with the mask applied in registers.
#include "attention_tc.cuh"
#include <stdlib.h>
#include <type_traits>

namespace mmvqa {

// Stage NT row-major [Tn, d] tiles (row strides differ) into shared memory as fp32.  VEC: 128-bit global loads,
// every thread issues all of its loads (up to 4 per round) before the first shared-memory store, so the CTA pays
// ONE global-memory latency for all tiles instead of one per element.
struct TileSrc {
  const void* src;
  int64_t row_stride;
  float* dst;
  int dst_ld;
};

template <typename T, int NT>
__device__ __forceinline__ void load_tiles(const TileSrc (&t)[NT], int Tn, int d, bool vec) {
  if (vec) {
    constexpr int N = Vec16<T>::N;
    const int cpr = d / N;                // 16-byte chunks per row
    const int per_tile = Tn * cpr;
    const int total = NT * per_tile;
    for (int base = threadIdx.x; base < total; base += blockDim.x * 4) {
      Vec16<T> v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = base + k * blockDim.x;
        if (idx < total) {
          const int ti = idx / per_tile, rem = idx - ti * per_tile;
          const int row = rem / cpr, ch = rem - row * cpr;
          v[k].load(reinterpret_cast<const T*>(t[ti].src) + row * t[ti].row_stride + ch * N);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int idx = base + k * blockDim.x;
        if (idx < total) {
          const int ti = idx / per_tile, rem = idx - ti * per_tile;
          const int row = rem / cpr, ch = rem - row * cpr;
          float* o = t[ti].dst + row * t[ti].dst_ld + ch * N;
#pragma unroll
          for (int j = 0; j < N; ++j) o[j] = v[k].get(j);
        }
      }
    }
  } else {
    for (int ti = 0; ti < NT; ++ti) {
      const T* src = reinterpret_cast<const T*>(t[ti].src);
      for (int idx = threadIdx.x; idx < Tn * d; idx += blockDim.x) {
        const int row = idx / d, c = idx - row * d;
        t[ti].dst[row * t[ti].dst_ld + c] = to_f(src[row * t[ti].row_stride + c]);
      }
    }
  }
}

// dot product of two shared-memory rows with four independent accumulators (breaks the FMA dependency chain)
__device__ __forceinline__ float dot_rows(const float* __restrict__ a, const float* __restrict__ b, int n) {
  float d0 = 0.0f, d1 = 0.0f, d2 = 0.0f, d3 = 0.0f;
  int s = 0;
#pragma unroll 2
  for (; s + 3 < n; s += 4) {
    d0 = fmaf(a[s], b[s], d0);
    d1 = fmaf(a[s + 1], b[s + 1], d1);
    d2 = fmaf(a[s + 2], b[s + 2], d2);
    d3 = fmaf(a[s + 3], b[s + 3], d3);
  }
  for (; s < n; ++s) d0 = fmaf(a[s], b[s], d0);
  return (d0 + d1) + (d2 + d3);
}
// sum_k a[k * lda] * b[k * ldb], two accumulators
__device__ __forceinline__ float dot_strided(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb, int n) {
  float d0 = 0.0f, d1 = 0.0f;
  int k = 0;
#pragma unroll 2
  for (; k + 1 < n; k += 2) {
    d0 = fmaf(a[k * lda], b[k * ldb], d0);
    d1 = fmaf(a[(k + 1) * lda], b[(k + 1) * ldb], d1);
  }
  if (k < n) d0 = fmaf(a[k * lda], b[k * ldb], d0);
  return d0 + d1;
}

template <typename T, bool RF>
__global__ void __launch_bounds__(1024) attn_fwd_kernel(const T* __restrict__ qkv, AttnLayout L,
                                                       const float* __restrict__ prev, const float* __restrict__ mask,
                                                       T* __restrict__ out, float* __restrict__ scores,
                                                       T* __restrict__ probs, int Tn, int heads, int d, float drop_p,
                                                       unsigned long long seed, const unsigned long long* seed_ctr, int vec) {
  extern __shared__ float sm[];
  const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Ks = sm;                       // [Tn][d+1]
  float* Qs = Ks + Tn * (d + 1);        // [Tn][d]
  float* Vs = Qs + Tn * d;              // [Tn][d]
  float* Ps = Vs + Tn * d;              // [nwarp][Tn]
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const T* base = qkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  pdl_wait();
  if (drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  pdl_trigger();
  {
    const TileSrc tiles[3] = {{base + L.k_off, L.row_stride, Ks, d + 1}, {base + L.q_off, L.row_stride, Qs, d},
                              {base + L.v_off, L.row_stride, Vs, d}};
    load_tiles<T, 3>(tiles, Tn, d, vec != 0);
  }
  __syncthreads();
  const float sqrt_d = sqrtf((float)d);
  const int64_t sbase = ((int64_t)b * heads + h) * Tn * Tn;
  const int H = heads * d;
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  // per-row global operands (prev row, query-mask term) are fetched one row ahead; the key-mask term of the
  // MHSA variant is row independent and fetched once
  float pn[4] = {0.0f, 0.0f, 0.0f, 0.0f}, qn = 0.0f, km[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  auto fetch = [&](int i) {
    if (RF) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = lane + 32 * jj;
        pn[jj] = (prev && j < Tn) ? __ldg(prev + sbase + (int64_t)i * Tn + j) : 0.0f;
      }
      qn = mask ? -10000.0f * (1.0f - __ldg(mask + b * Tn + i)) : 0.0f;
    }
  };
  if (!RF && mask) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) km[jj] = (lane + 32 * jj < Tn) ? -10000.0f * (1.0f - __ldg(mask + b * Tn + lane + 32 * jj)) : 0.0f;
  }
  if (warp < Tn) fetch(warp);
  for (int i = warp; i < Tn; i += nwarp) {
    float sc[4];
    float pc[4] = {pn[0], pn[1], pn[2], pn[3]};
    const float qoff = qn;
    if (i + nwarp < Tn) fetch(i + nwarp);
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      sc[jj] = -INFINITY;
      if (j < Tn) {
        const float dot = dot_rows(Qs + i * d, Ks + j * (d + 1), d);
        float v = dot / sqrt_d;
        if (RF) {
          v += pc[jj];
          v += qoff;
          scores[sbase + (int64_t)i * Tn + j] = v;
        } else {
          v += km[jj];
        }
        sc[jj] = v;
        mx = fmaxf(mx, v);
      }
    }
    mx = warp_max(mx);
    float sum = 0.0f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      if (j < Tn) {
        sc[jj] = expf(sc[jj] - mx);
        sum += sc[jj];
      }
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      if (j < Tn) {
        float pr = sc[jj] * inv;
        if (!RF) {
          probs[sbase + (int64_t)i * Tn + j] = from_f<T>(pr);
          if (drop_p > 0.0f) pr = hash32(seed, (uint64_t)(sbase + (int64_t)i * Tn + j)) >= thr ? pr * inv_keep : 0.0f;
        }
        Ps[warp * Tn + j] = pr;
      }
    }
    __syncwarp();
    for (int s = lane; s < d; s += 32) {
      const float acc = dot_strided(Ps + warp * Tn, 1, Vs + s, d, Tn);
      out[((int64_t)b * Tn + i) * H + h * d + s] = from_f<T>(acc);
    }
    __syncwarp();
  }
}

template <typename T, bool RF>
__global__ void __launch_bounds__(1024) attn_bwd_kernel(const T* __restrict__ qkv, AttnLayout L,
                                                       const float* __restrict__ scores, const T* __restrict__ probs,
                                                       const T* __restrict__ dout, const float* __restrict__ dscores_in,
                                                       T* __restrict__ dqkv, float* __restrict__ dprev, int Tn, int heads,
                                                       int d, float drop_p, unsigned long long seed, const unsigned long long* seed_ctr, int vec,
                                                       int resident) {
  extern __shared__ float sm[];
  const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* X = sm;                        // [Tn][d+1]  V (then K, then Q when the tiles do not all fit)
  float* dO = X + Tn * (d + 1);         // [Tn][d]
  float* P = dO + Tn * d;               // [Tn][Tn]   probabilities used by dV (post-dropout for MHSA)
  float* dS = P + Tn * Tn;              // [Tn][Tn]   gradient of the pre-softmax scores
  float* Kb = resident ? dS + Tn * Tn : X;          // resident: K and Q have their own tiles and are
  float* Qb = resident ? Kb + Tn * (d + 1) : X;     // fetched together with V and dO (one latency)
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int H = heads * d;
  const T* base = qkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  T* dbase = dqkv + (int64_t)b * L.tok_batch + (int64_t)h * L.head_stride;
  const int64_t sbase = ((int64_t)b * heads + h) * Tn * Tn;
  pdl_wait();
  if (drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  pdl_trigger();
  if (resident) {
    const TileSrc tiles[4] = {{base + L.v_off, L.row_stride, X, d + 1}, {dout + (int64_t)b * Tn * H + h * d, (int64_t)H, dO, d},
                              {base + L.k_off, L.row_stride, Kb, d + 1}, {base + L.q_off, L.row_stride, Qb, d + 1}};
    load_tiles<T, 4>(tiles, Tn, d, vec != 0);
  } else {
    const TileSrc tiles[2] = {{base + L.v_off, L.row_stride, X, d + 1}, {dout + (int64_t)b * Tn * H + h * d, (int64_t)H, dO, d}};
    load_tiles<T, 2>(tiles, Tn, d, vec != 0);
  }
  __syncthreads();
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  // score / probability rows and the incoming score gradient are fetched one row ahead
  float sn[4], gn[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  auto fetch = [&](int i) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      sn[jj] = RF ? -INFINITY : 0.0f;
      gn[jj] = 0.0f;
      if (j < Tn) {
        if (RF) {
          sn[jj] = __ldg(scores + sbase + (int64_t)i * Tn + j);
          if (dscores_in) gn[jj] = __ldg(dscores_in + sbase + (int64_t)i * Tn + j);
        } else {
          sn[jj] = to_f(probs[sbase + (int64_t)i * Tn + j]);
        }
      }
    }
  };
  if (warp < Tn) fetch(warp);
  for (int i = warp; i < Tn; i += nwarp) {
    float pr[4], dp[4], gin[4];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      pr[jj] = sn[jj];
      gin[jj] = gn[jj];
      if (RF) mx = fmaxf(mx, pr[jj]);
    }
    if (i + nwarp < Tn) fetch(i + nwarp);
    if (RF) {
      mx = warp_max(mx);
      float sum = 0.0f;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj)
        if (lane + 32 * jj < Tn) {
          pr[jj] = expf(pr[jj] - mx);
          sum += pr[jj];
        }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) pr[jj] = (lane + 32 * jj < Tn) ? pr[jj] * inv : 0.0f;
    }
    float rowdot = 0.0f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      dp[jj] = 0.0f;
      if (j < Tn) {
        float acc = dot_rows(dO + i * d, X + j * (d + 1), d);
        float pd = pr[jj];
        if (!RF && drop_p > 0.0f) {
          const bool keep = hash32(seed, (uint64_t)(sbase + (int64_t)i * Tn + j)) >= thr;
          acc = keep ? acc * inv_keep : 0.0f;
          pd = keep ? pd * inv_keep : 0.0f;
        }
        P[i * Tn + j] = pd;
        dp[jj] = acc;
        rowdot += pr[jj] * acc;
      }
    }
    rowdot = warp_sum(rowdot);
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = lane + 32 * jj;
      if (j < Tn) {
        float g = pr[jj] * (dp[jj] - rowdot);
        if (RF) {
          g += gin[jj];
          if (dprev) dprev[sbase + (int64_t)i * Tn + j] = g;
        }
        dS[i * Tn + j] = g;
      }
    }
  }
  __syncthreads();
  const float inv_sqrt_d = 1.0f / sqrtf((float)d);
  // dV[j][s] = sum_i P[i][j] dO[i][s]
  for (int idx = threadIdx.x; idx < Tn * d; idx += blockDim.x) {
    const int j = idx / d, s = idx - j * d;
    const float acc = dot_strided(P + j, Tn, dO + s, d, Tn);
    dbase[L.v_off + (int64_t)j * L.row_stride + s] = from_f<T>(acc);
  }
  if (!resident) {
    __syncthreads();
    const TileSrc tk[1] = {{base + L.k_off, L.row_stride, X, d + 1}};
    load_tiles<T, 1>(tk, Tn, d, vec != 0);
    __syncthreads();
  }
  // dQ[i][s] = sum_j dS[i][j] K[j][s] / sqrt(d)
  for (int idx = threadIdx.x; idx < Tn * d; idx += blockDim.x) {
    const int i = idx / d, s = idx - i * d;
    const float acc = dot_strided(dS + i * Tn, 1, Kb + s, d + 1, Tn);
    dbase[L.q_off + (int64_t)i * L.row_stride + s] = from_f<T>(acc * inv_sqrt_d);
  }
  if (!resident) {
    __syncthreads();
    const TileSrc tq[1] = {{base + L.q_off, L.row_stride, X, d + 1}};
    load_tiles<T, 1>(tq, Tn, d, vec != 0);
    __syncthreads();
  }
  // dK[j][s] = sum_i dS[i][j] Q[i][s] / sqrt(d)
  for (int idx = threadIdx.x; idx < Tn * d; idx += blockDim.x) {
    const int j = idx / d, s = idx - j * d;
    const float acc = dot_strided(dS + j, Tn, Qb + s, d + 1, Tn);
    dbase[L.k_off + (int64_t)j * L.row_stride + s] = from_f<T>(acc * inv_sqrt_d);
  }
}

// the tensor-core variant needs 16-byte aligned rows (d, strides and offsets multiples of 8 bf16) and d % 16 == 0
static bool tc_attention_ok(const void* qkv, const AttnLayout& L, const void* dout, int Tn, int heads, int d) {
  static const bool off = getenv("MMVQA_NO_TC_ATTN") != nullptr;
  if (off || d % 16 != 0 || d > 128 || Tn > 128) return false;
  if (L.row_stride % 8 || L.head_stride % 8 || L.q_off % 8 || L.k_off % 8 || L.v_off % 8) return false;
  if (reinterpret_cast<uintptr_t>(qkv) & 15) return false;
  if (dout && ((reinterpret_cast<uintptr_t>(dout) & 15) || (heads * d) % 8)) return false;
  return true;
}

static int check_shape(const char* name, int B, int T, int heads, int d) {
  MMVQA_REQUIRE(B > 0 && T > 0 && heads > 0 && d > 0, "%s: bad shape", name);
  MMVQA_REQUIRE(T <= 128, "%s: sequence length %d > 128 (short-sequence kernel)", name, T);
  MMVQA_REQUIRE(d <= 128, "%s: head dim %d > 128", name, d);
  return MMVQA_OK;
}

template <typename T, bool RF>
static int launch_fwd(const void* qkv, const AttnLayout& L, const float* prev, const float* mask, void* out, float* scores,
                      void* probs, int B, int Tn, int heads, int d, float p, uint64_t seed, cudaStream_t st) {
  if (std::is_same<T, __nv_bfloat16>::value && tc_attention_ok(qkv, L, nullptr, Tn, heads, d)) {
    const int Tp = (Tn + 15) & ~15, ldn = d + 8, ldt = Tp + 8;
    const size_t smem_tc = (size_t)(2 * Tp * ldn + d * ldt) * 2;
    if (smem_tc <= 227 * 1024) {
      const int nthr = 32 * (Tp / 16 < 4 ? 4 : Tp / 16);
#define TC_FWD(TPV)                                                                                                     \
  do {                                                                                                                  \
    auto k = attn_tc_fwd_kernel<RF, TPV>;                                                                               \
    if (smem_tc > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc)); \
    MMVQA_CUDA(launch_pdl(k, dim3(B * heads), dim3(nthr), smem_tc, st, (const bf16*)qkv, L, prev, mask, (bf16*)out, scores, \
                          (bf16*)probs, Tn, heads, d, p, (unsigned long long)seed, g_seed_ctr, (const bf16*)nullptr,    \
                          (const bf16*)nullptr));                                                                       \
  } while (0)
      if (Tp <= 32) TC_FWD(32); else if (Tp <= 64) TC_FWD(64); else if (Tp <= 96) TC_FWD(96); else TC_FWD(128);
#undef TC_FWD
      MMVQA_LAUNCHED("attn_tc_fwd");
      return MMVQA_OK;
    }
  }
  const int nthreads = 32 * (Tn < 8 ? 8 : (Tn > 32 ? 32 : Tn));   // one warp per query row, 8..32 warps
  size_t smem = sizeof(float) * ((size_t)Tn * (d + 1) + 2 * (size_t)Tn * d + (size_t)(nthreads / 32) * Tn);
  if (smem > 227 * 1024) return set_err(MMVQA_ERR_SMEM, "attention fwd: T=%d d=%d needs %zu bytes of shared memory", Tn, d, smem);
  auto kern = attn_fwd_kernel<T, RF>;
  if (smem > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vn = 16 / (int)sizeof(T);
  const int vec = (d % vn == 0 && L.row_stride % vn == 0 && L.head_stride % vn == 0 && L.q_off % vn == 0 && L.k_off % vn == 0 &&
                   L.v_off % vn == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0) ? 1 : 0;
  MMVQA_CUDA(launch_pdl(kern, dim3(B * heads), dim3(nthreads), smem, st, (const T*)qkv, L, prev, mask, (T*)out, scores, (T*)probs, Tn, heads, d, p,
                        (unsigned long long)seed, g_seed_ctr, vec));
  MMVQA_LAUNCHED("attn_fwd");
  return MMVQA_OK;
}

template <typename T, bool RF>
static int launch_bwd(const void* qkv, const AttnLayout& L, const float* scores, const void* probs, const void* dout,
                      const float* dscores_in, void* dqkv, float* dprev, int B, int Tn, int heads, int d, float p,
                      uint64_t seed, cudaStream_t st) {
  if (std::is_same<T, __nv_bfloat16>::value && tc_attention_ok(qkv, L, dout, Tn, heads, d) &&
      (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0) {
    const int Tp = (Tn + 15) & ~15, ldn = d + 8, ldt = Tp + 8;
    const size_t smem_tc = (size_t)(2 * Tp * ldn + 3 * d * ldt + 2 * Tp * ldt) * 2;
    if (smem_tc <= 227 * 1024) {
      // one warp per 16-row strip and no fewer than `minw`: 204 registers per thread make a 4-warp CTA a 26.6 K-register
      // tenant that cannot share an SM with two resident GEMM CTAs (2 x 30.7 K of 64 K)
      static const int minw = getenv("MMVQA_ATTN_BWD_WARPS") ? atoi(getenv("MMVQA_ATTN_BWD_WARPS")) : 4;
      const int nthr = 32 * (Tp / 16 < minw ? minw : Tp / 16);
#define TC_BWD(TPV)                                                                                                     \
  do {                                                                                                                  \
    auto k = attn_tc_bwd_kernel<RF, TPV>;                                                                               \
    if (smem_tc > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tc)); \
    MMVQA_CUDA(launch_pdl(k, dim3(B * heads), dim3(nthr), smem_tc, st, (const bf16*)qkv, L, scores, (const bf16*)probs,  \
                          (const bf16*)dout, dscores_in, (bf16*)dqkv, dprev, Tn, heads, d, p, (unsigned long long)seed, \
                          g_seed_ctr, (const bf16*)nullptr, (const bf16*)nullptr, (bf16*)nullptr));                     \
  } while (0)
      if (Tp <= 32) TC_BWD(32); else if (Tp <= 64) TC_BWD(64); else if (Tp <= 96) TC_BWD(96); else TC_BWD(128);
#undef TC_BWD
      MMVQA_LAUNCHED("attn_tc_bwd");
      return MMVQA_OK;
    }
  }
  size_t smem = sizeof(float) * ((size_t)Tn * (d + 1) + (size_t)Tn * d + 2 * (size_t)Tn * Tn);
  const size_t smem_res = smem + sizeof(float) * 2 * (size_t)Tn * (d + 1);
  const int resident = smem_res <= 100 * 1024 ? 1 : 0;   // K and Q tiles too when two CTAs still share an SM
  if (resident) smem = smem_res;
  if (smem > 227 * 1024) return set_err(MMVQA_ERR_SMEM, "attention bwd: T=%d d=%d needs %zu bytes of shared memory", Tn, d, smem);
  auto kern = attn_bwd_kernel<T, RF>;
  if (smem > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int vn = 16 / (int)sizeof(T);
  const int H = heads * d;
  const int vec = (d % vn == 0 && L.row_stride % vn == 0 && L.head_stride % vn == 0 && L.q_off % vn == 0 && L.k_off % vn == 0 &&
                   L.v_off % vn == 0 && H % vn == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dout) & 15) == 0) ? 1 : 0;
  const int nthreads = 32 * (Tn < 8 ? 8 : (Tn > 32 ? 32 : Tn));
  MMVQA_CUDA(launch_pdl(kern, dim3(B * heads), dim3(nthreads), smem, st, (const T*)qkv, L, scores, (const T*)probs, (const T*)dout,
                        dscores_in, (T*)dqkv, dprev, Tn, heads, d, p, (unsigned long long)seed, g_seed_ctr, vec, resident));
  MMVQA_LAUNCHED("attn_bwd");
  return MMVQA_OK;
}

static AttnLayout rf_layout(int T, int heads, int d) {
  AttnLayout L;
  L.row_stride = (int64_t)heads * 3 * d;
  L.head_stride = 3 * d;
  L.tok_batch = (int64_t)T * L.row_stride;
  L.k_off = 0; L.q_off = d; L.v_off = 2 * d;     // split order k, q, v (realformer.py:33)
  return L;
}
static AttnLayout mhsa_layout(int T, int heads, int d) {
  AttnLayout L;
  const int H = heads * d;
  L.row_stride = 3 * (int64_t)H;
  L.head_stride = d;
  L.tok_batch = (int64_t)T * L.row_stride;
  L.q_off = 0; L.k_off = H; L.v_off = 2 * H;     // fused [q | k | v] projection
  return L;
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_rf_attn_fwd(const void* kqv, const float* prev, const float* mask, void* out, float* scores, int B, int T,
                      int heads, int d, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(kqv && out && scores, "rf_attn_fwd: null pointer");
  int rc = check_shape("rf_attn_fwd", B, T, heads, d);
  if (rc) return rc;
  AttnLayout L = rf_layout(T, heads, d);
  if (dtype == MMVQA_F32) return launch_fwd<float, true>(kqv, L, prev, mask, out, scores, nullptr, B, T, heads, d, 0.f, 0, as_stream(stream));
  if (dtype == MMVQA_BF16) return launch_fwd<__nv_bfloat16, true>(kqv, L, prev, mask, out, scores, nullptr, B, T, heads, d, 0.f, 0, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "rf_attn_fwd: bad dtype %d", dtype);
}

// RealFormer attention with the kqv projection inside (bf16 tensor-core path only): x [B*T, heads*d], wkqv [3d, d]
int mmvqa_rf_attn_fwd_fused(const void* x, const void* wkqv, const float* prev, const float* mask, void* out, float* scores,
                            void* kqv_out, int B, int T, int heads, int d, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(x && wkqv && out && scores && kqv_out, "rf_attn_fwd_fused: null pointer");
  int rc = check_shape("rf_attn_fwd_fused", B, T, heads, d);
  if (rc) return rc;
  MMVQA_REQUIRE(dtype == MMVQA_BF16, "rf_attn_fwd_fused: bf16 only (the fp32 path keeps the separate kqv GEMM)");
  AttnLayout L = rf_layout(T, heads, d);
  MMVQA_REQUIRE(tc_attention_ok(kqv_out, L, nullptr, T, heads, d) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wkqv)) & 15) == 0,
                "rf_attn_fwd_fused: needs d %% 16 == 0, T <= 128 and 16-byte aligned operands");
  const int Tp = (T + 15) & ~15, ldn = d + 8, ldt = Tp + 8;
  const size_t smem = (size_t)(2 * Tp * ldn + d * ldt + Tp * ldn + 3 * d * ldn) * 2;
  MMVQA_REQUIRE(smem <= 227 * 1024, "rf_attn_fwd_fused: T=%d d=%d needs %zu bytes of shared memory", T, d, smem);
  const int nthr = 256;
  cudaStream_t st = as_stream(stream);
#define TC_FWDF(TPV)                                                                                                    \
  do {                                                                                                                  \
    auto k = attn_tc_fwd_kernel<true, TPV, true>;                                                                       \
    if (smem > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MMVQA_CUDA(launch_pdl(k, dim3(B * heads), dim3(nthr), smem, st, (const bf16*)kqv_out, L, prev, mask, (bf16*)out, scores, \
                          (bf16*)nullptr, T, heads, d, 0.0f, 0ull, g_seed_ctr, (const bf16*)x, (const bf16*)wkqv));   \
  } while (0)
  if (Tp <= 32) TC_FWDF(32); else if (Tp <= 64) TC_FWDF(64); else if (Tp <= 96) TC_FWDF(96); else TC_FWDF(128);
#undef TC_FWDF
  MMVQA_LAUNCHED("rf_attn_fwd_fused");
  return MMVQA_OK;
}

int mmvqa_rf_attn_bwd(const void* kqv, const float* scores, const void* dout, const float* dscores_in, void* dkqv,
                      float* dprev, int B, int T, int heads, int d, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(kqv && scores && dout && dkqv, "rf_attn_bwd: null pointer");
  int rc = check_shape("rf_attn_bwd", B, T, heads, d);
  if (rc) return rc;
  AttnLayout L = rf_layout(T, heads, d);
  if (dtype == MMVQA_F32) return launch_bwd<float, true>(kqv, L, scores, nullptr, dout, dscores_in, dkqv, dprev, B, T, heads, d, 0.f, 0, as_stream(stream));
  if (dtype == MMVQA_BF16) return launch_bwd<__nv_bfloat16, true>(kqv, L, scores, nullptr, dout, dscores_in, dkqv, dprev, B, T, heads, d, 0.f, 0, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "rf_attn_bwd: bad dtype %d", dtype);
}

// RealFormer attention backward with the kqv input gradient inside (bf16 tensor-core path only)
int mmvqa_rf_attn_bwd_fused(const void* kqv, const float* scores, const void* dout, const float* dscores_in, void* dkqv,
                            float* dprev, const void* wkqv, const void* dres, void* dx, int B, int T, int heads, int d,
                            int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(kqv && scores && dout && dkqv && wkqv && dx, "rf_attn_bwd_fused: null pointer");
  int rc = check_shape("rf_attn_bwd_fused", B, T, heads, d);
  if (rc) return rc;
  MMVQA_REQUIRE(dtype == MMVQA_BF16, "rf_attn_bwd_fused: bf16 only (the fp32 path keeps the separate dgrad GEMM)");
  AttnLayout L = rf_layout(T, heads, d);
  MMVQA_REQUIRE(tc_attention_ok(kqv, L, dout, T, heads, d) && (reinterpret_cast<uintptr_t>(dkqv) & 15) == 0 &&
                    ((reinterpret_cast<uintptr_t>(wkqv) | reinterpret_cast<uintptr_t>(dx) | reinterpret_cast<uintptr_t>(dres)) & 15) == 0,
                "rf_attn_bwd_fused: needs d %% 16 == 0, T <= 128 and 16-byte aligned operands");
  const int Tp = (T + 15) & ~15, ldn = d + 8, ldt = Tp + 8, ldg = 3 * d + 8;
  const size_t smem = (size_t)(2 * Tp * ldn + 3 * d * ldt + 2 * Tp * ldt + Tp * ldg + 3 * d * ldn) * 2;
  MMVQA_REQUIRE(smem <= 227 * 1024, "rf_attn_bwd_fused: T=%d d=%d needs %zu bytes of shared memory", T, d, smem);
  cudaStream_t st = as_stream(stream);
#define TC_BWDF(TPV)                                                                                                    \
  do {                                                                                                                  \
    auto k = attn_tc_bwd_kernel<true, TPV, true>;                                                                       \
    if (smem > 48 * 1024) MMVQA_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    MMVQA_CUDA(launch_pdl(k, dim3(B * heads), dim3(256), smem, st, (const bf16*)kqv, L, scores, (const bf16*)nullptr,    \
                          (const bf16*)dout, dscores_in, (bf16*)dkqv, dprev, T, heads, d, 0.0f, 0ull, g_seed_ctr,       \
                          (const bf16*)wkqv, (const bf16*)dres, (bf16*)dx));                                            \
  } while (0)
  if (Tp <= 32) TC_BWDF(32); else if (Tp <= 64) TC_BWDF(64); else if (Tp <= 96) TC_BWDF(96); else TC_BWDF(128);
#undef TC_BWDF
  MMVQA_LAUNCHED("rf_attn_bwd_fused");
  return MMVQA_OK;
}

int mmvqa_mhsa_fwd(const void* qkv, const float* mask, void* out, void* probs, int B, int T, int heads, int d,
                   float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(qkv && out && probs, "mhsa_fwd: null pointer");
  MMVQA_REQUIRE(dropout_p >= 0.0f && dropout_p < 1.0f, "mhsa_fwd: dropout_p must be in [0,1)");
  int rc = check_shape("mhsa_fwd", B, T, heads, d);
  if (rc) return rc;
  AttnLayout L = mhsa_layout(T, heads, d);
  if (dtype == MMVQA_F32) return launch_fwd<float, false>(qkv, L, nullptr, mask, out, nullptr, probs, B, T, heads, d, dropout_p, dropout_seed, as_stream(stream));
  if (dtype == MMVQA_BF16) return launch_fwd<__nv_bfloat16, false>(qkv, L, nullptr, mask, out, nullptr, probs, B, T, heads, d, dropout_p, dropout_seed, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "mhsa_fwd: bad dtype %d", dtype);
}

int mmvqa_mhsa_bwd(const void* qkv, const void* probs, const void* dout, void* dqkv, int B, int T, int heads, int d,
                   float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(qkv && probs && dout && dqkv, "mhsa_bwd: null pointer");
  int rc = check_shape("mhsa_bwd", B, T, heads, d);
  if (rc) return rc;
  AttnLayout L = mhsa_layout(T, heads, d);
  if (dtype == MMVQA_F32) return launch_bwd<float, false>(qkv, L, nullptr, probs, dout, nullptr, dqkv, nullptr, B, T, heads, d, dropout_p, dropout_seed, as_stream(stream));
  if (dtype == MMVQA_BF16) return launch_bwd<__nv_bfloat16, false>(qkv, L, nullptr, probs, dout, nullptr, dqkv, nullptr, B, T, heads, d, dropout_p, dropout_seed, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "mhsa_bwd: bad dtype %d", dtype);
}

}  // extern "C"
