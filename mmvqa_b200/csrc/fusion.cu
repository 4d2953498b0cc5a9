// fusion.cu -- input fusion and pooling kernels of the MMBERT path:
//   * BertEmbeddings (word + position + token-type -> LayerNorm -> dropout) fused with the
//     visual-token overwrite of positions 0..nvis-1       (models/mmbert.py:60-67)
//   * mask-weighted mean pooling                           (models/mmbert.py:169-172)
//   * row-wise L2 normalisation                            (models/mmbert.py:157)
#include "common.cuh"

namespace mmvqa {

// one warp per token row (b, t)
template <typename T>
__global__ void __launch_bounds__(128) embed_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ seg,
                                                        const float* __restrict__ word, const float* __restrict__ pos,
                                                        const float* __restrict__ typ, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ vis,
                                                        T* __restrict__ h, float* __restrict__ mean_out,
                                                        float* __restrict__ rstd_out, int B, int Tn, int H, int nvis,
                                                        float eps, float drop_p, unsigned long long seed,
                                                        const unsigned long long* seed_ctr) {
  if (drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (int64_t)B * Tn) return;
  const int b = (int)(row / Tn), t = (int)(row % Tn);
  T* hr = h + row * H;
  if (t < nvis) {  // visual token n = t of sample b
    const float* v = vis + ((int64_t)t * B + b) * H;
    for (int c = lane; c < H; c += 32) hr[c] = from_f<T>(v[c]);
    if (lane == 0) {
      mean_out[row] = 0.0f;
      rstd_out[row] = 0.0f;
    }
    return;
  }
  const float* w = word + ids[row] * (int64_t)H;
  const float* p = pos + (int64_t)t * H;
  const float* ty = typ + seg[row] * (int64_t)H;
  float s = 0.0f;
  for (int c = lane; c < H; c += 32) s += w[c] + ty[c] + p[c];
  const float mean = warp_sum(s) / (float)H;
  float q = 0.0f;
  for (int c = lane; c < H; c += 32) {
    float d = (w[c] + ty[c] + p[c]) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  for (int c = lane; c < H; c += 32) {
    float y = ((w[c] + ty[c] + p[c]) - mean) * rstd * gamma[c] + beta[c];
    if (drop_p > 0.0f) y = hash32(seed, (uint64_t)(row * H + c)) >= thr ? y * inv_keep : 0.0f;
    hr[c] = from_f<T>(y);
  }
}

template <typename T>
__global__ void __launch_bounds__(128) embed_bwd_kernel(const T* __restrict__ dh, const int64_t* __restrict__ ids,
                                                        const int64_t* __restrict__ seg, const float* __restrict__ word,
                                                        const float* __restrict__ pos, const float* __restrict__ typ,
                                                        const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, float* __restrict__ dword,
                                                        float* __restrict__ dpos, float* __restrict__ dtyp,
                                                        float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                        float* __restrict__ dvis, int B, int Tn, int H, int nvis,
                                                        int padding_idx, float drop_p, unsigned long long seed,
                                                        const unsigned long long* seed_ctr) {
  if (drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  extern __shared__ float sm[];  // [2][H] dgamma / dbeta block partials
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) sm[c] = 0.0f;
  __syncthreads();
  const int64_t rows = (int64_t)B * Tn;
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  for (int64_t row = blockIdx.x * (int64_t)nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    const int b = (int)(row / Tn), t = (int)(row % Tn);
    const T* dr = dh + row * H;
    if (t < nvis) {
      if (dvis) {
        float* dv = dvis + ((int64_t)t * B + b) * H;
        for (int c = lane; c < H; c += 32) dv[c] = to_f(dr[c]);
      }
      continue;
    }
    const int64_t id = ids[row], sg = seg[row];
    const float* w = word + id * (int64_t)H;
    const float* p = pos + (int64_t)t * H;
    const float* ty = typ + sg * (int64_t)H;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.0f, s2 = 0.0f;
    for (int c = lane; c < H; c += 32) {
      float g = to_f(dr[c]);
      if (drop_p > 0.0f) g = hash32(seed, (uint64_t)(row * H + c)) >= thr ? g * inv_keep : 0.0f;
      float xh = ((w[c] + ty[c] + p[c]) - mu) * rs;
      atomicAdd(&sm[c], g * xh);
      atomicAdd(&sm[H + c], g);
      g *= gamma[c];
      s1 += g;
      s2 += g * xh;
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    for (int c = lane; c < H; c += 32) {
      float g = to_f(dr[c]);
      if (drop_p > 0.0f) g = hash32(seed, (uint64_t)(row * H + c)) >= thr ? g * inv_keep : 0.0f;
      g *= gamma[c];
      float xh = ((w[c] + ty[c] + p[c]) - mu) * rs;
      float de = rs * (g - s1 - xh * s2);
      if (dword && id != padding_idx) atomicAdd(dword + id * (int64_t)H + c, de);
      if (dpos) atomicAdd(dpos + (int64_t)t * H + c, de);
      if (dtyp) atomicAdd(dtyp + sg * (int64_t)H + c, de);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, sm[c]);
    if (dbeta) atomicAdd(dbeta + c, sm[H + c]);
  }
}

// Vector variant (bf16 gradient, H = KU * 256): a lane owns 8 consecutive columns of each of its KU slots, so the LayerNorm
// gamma / beta gradients and the token-type rows 0 / 1 (every token of the batch lands on one of those two rows: B*T-way
// contention when done with global atomics) accumulate in registers over the warp's rows and leave the CTA once, through
// a shared-memory slab; word / position rows go out as 128-bit reductions.
template <int KU>
__global__ void __launch_bounds__(256) embed_bwd_vec_kernel(
    const __nv_bfloat16* __restrict__ dh, const int64_t* __restrict__ ids, const int64_t* __restrict__ seg,
    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ typ,
    const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
    float* __restrict__ dword, float* __restrict__ dpos, float* __restrict__ dtyp, float* __restrict__ dgamma,
    float* __restrict__ dbeta, float* __restrict__ dvis, int B, int Tn, int nvis, int padding_idx, float drop_p,
    unsigned long long seed, const unsigned long long* seed_ctr) {
  constexpr int H = KU * 256;
  if (drop_p > 0.0f) seed = seed_eff(seed, seed_ctr);
  __shared__ __align__(16) float sm[4 * H];   // dgamma | dbeta | dtyp row 0 | dtyp row 1
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int c = threadIdx.x; c < 4 * H; c += blockDim.x) sm[c] = 0.0f;
  float dg[KU * 8], db[KU * 8], dt0[KU * 8], dt1[KU * 8];
#pragma unroll
  for (int k = 0; k < KU * 8; ++k) dg[k] = db[k] = dt0[k] = dt1[k] = 0.0f;
  const int64_t rows = (int64_t)B * Tn;
  const uint32_t thr = (uint32_t)(drop_p * 4294967296.0);
  const float inv_keep = drop_p > 0.0f ? 1.0f / (1.0f - drop_p) : 1.0f;
  for (int64_t row = blockIdx.x * (int64_t)nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    const int b = (int)(row / Tn), t = (int)(row % Tn);
    const __nv_bfloat16* dr = dh + row * H;
    if (t < nvis) {
      if (dvis) {
        float* dv = dvis + ((int64_t)t * B + b) * H;
#pragma unroll
        for (int k = 0; k < KU; ++k) {
          const int c = (k * 32 + lane) * 8;
          Vec16<__nv_bfloat16> v;
          v.load(dr + c);
          *reinterpret_cast<float4*>(dv + c) = make_float4(v.get(0), v.get(1), v.get(2), v.get(3));
          *reinterpret_cast<float4*>(dv + c + 4) = make_float4(v.get(4), v.get(5), v.get(6), v.get(7));
        }
      }
      continue;
    }
    const int64_t id = ids[row], sg = seg[row];
    const float* w = word + id * (int64_t)H;
    const float* p = pos + (int64_t)t * H;
    const float* ty = typ + sg * (int64_t)H;
    const float mu = mean[row], rs = rstd[row];
    float xh[KU * 8], gg[KU * 8];
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      const int c = (k * 32 + lane) * 8;
      Vec16<__nv_bfloat16> v;
      v.load(dr + c);
#pragma unroll
      for (int h4 = 0; h4 < 2; ++h4) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + c) + h4);
        const float4 pv = __ldg(reinterpret_cast<const float4*>(p + c) + h4);
        const float4 tv = __ldg(reinterpret_cast<const float4*>(ty + c) + h4);
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c) + h4);
        const float xs[4] = {(wv.x + tv.x) + pv.x, (wv.y + tv.y) + pv.y, (wv.z + tv.z) + pv.z, (wv.w + tv.w) + pv.w};
        const float gs[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int e = h4 * 4 + j, q = k * 8 + e;
          float g = v.get(e);
          if (drop_p > 0.0f) g = hash32(seed, (uint64_t)(row * H + c + e)) >= thr ? g * inv_keep : 0.0f;
          const float x = (xs[j] - mu) * rs;
          dg[q] += g * x;
          db[q] += g;
          g *= gs[j];
          xh[q] = x;
          gg[q] = g;
          s1 += g;
          s2 += g * x;
        }
      }
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    const bool to_word = dword != nullptr && id != padding_idx;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      const int c = (k * 32 + lane) * 8;
      float de[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) de[e] = rs * (gg[k * 8 + e] - s1 - xh[k * 8 + e] * s2);
      const float4 lo = make_float4(de[0], de[1], de[2], de[3]), hi = make_float4(de[4], de[5], de[6], de[7]);
      if (to_word) {
        atomicAdd(reinterpret_cast<float4*>(dword + id * (int64_t)H + c), lo);
        atomicAdd(reinterpret_cast<float4*>(dword + id * (int64_t)H + c) + 1, hi);
      }
      if (dpos) {
        atomicAdd(reinterpret_cast<float4*>(dpos + (int64_t)t * H + c), lo);
        atomicAdd(reinterpret_cast<float4*>(dpos + (int64_t)t * H + c) + 1, hi);
      }
      if (dtyp) {
        if (sg == 0) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dt0[k * 8 + e] += de[e];
        } else if (sg == 1) {
#pragma unroll
          for (int e = 0; e < 8; ++e) dt1[k * 8 + e] += de[e];
        } else {
          atomicAdd(reinterpret_cast<float4*>(dtyp + sg * (int64_t)H + c), lo);
          atomicAdd(reinterpret_cast<float4*>(dtyp + sg * (int64_t)H + c) + 1, hi);
        }
      }
    }
  }
  __syncthreads();   // slab zeroed by everyone
#pragma unroll
  for (int k = 0; k < KU; ++k) {
    const int c = (k * 32 + lane) * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&sm[c + e], dg[k * 8 + e]);
      atomicAdd(&sm[H + c + e], db[k * 8 + e]);
      atomicAdd(&sm[2 * H + c + e], dt0[k * 8 + e]);
      atomicAdd(&sm[3 * H + c + e], dt1[k * 8 + e]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {   // 4 * H / 4 vectors
    const int which = i / (H / 4), c4 = (i - which * (H / 4)) * 4;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : (dtyp ? dtyp + (which - 2) * (int64_t)H : nullptr));
    if (dst == nullptr) continue;
    const float4 v = *reinterpret_cast<const float4*>(sm + which * H + c4);
    if (which >= 2 && v.x == 0.0f && v.y == 0.0f && v.z == 0.0f && v.w == 0.0f) continue;   // row 1 may not exist
    atomicAdd(reinterpret_cast<float4*>(dst + c4), v);
  }
}

// out[b, c] = sum_t h[b,t,c] * mask[b,t] / max(sum_t mask[b,t], 1e-9)
template <typename T>
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const T* __restrict__ h, const float* __restrict__ mask,
                                                              T* __restrict__ out, int Tn, int H) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float den = 0.0f;
  for (int t = 0; t < Tn; ++t) den += mask[b * Tn + t];
  den = fmaxf(den, 1e-9f);
  if (c >= H) return;
  float acc = 0.0f;
  for (int t = 0; t < Tn; ++t) acc += to_f(h[((int64_t)b * Tn + t) * H + c]) * mask[b * Tn + t];
  out[(int64_t)b * H + c] = from_f<T>(acc / den);
}
template <typename T>
__global__ void __launch_bounds__(256) masked_mean_bwd_kernel(const T* __restrict__ dout, const float* __restrict__ mask,
                                                              T* __restrict__ dh, int Tn, int H) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  float den = 0.0f;
  for (int t = 0; t < Tn; ++t) den += mask[b * Tn + t];
  den = fmaxf(den, 1e-9f);
  if (c >= H) return;
  const float g = to_f(dout[(int64_t)b * H + c]) / den;
  for (int t = 0; t < Tn; ++t) dh[((int64_t)b * Tn + t) * H + c] = from_f<T>(g * mask[b * Tn + t]);
}

// y = x / max(||x||, 1e-12)   (F.normalize, dim=1); one warp per row
__global__ void __launch_bounds__(128) l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ inv_norm, int rows, int cols) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.0f;
  for (int c = lane; c < cols; c += 32) s += x[(int64_t)row * cols + c] * x[(int64_t)row * cols + c];
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int c = lane; c < cols; c += 32) y[(int64_t)row * cols + c] = x[(int64_t)row * cols + c] * inv;
}
// dx = (dy - y (y . dy)) * inv_norm
__global__ void __launch_bounds__(128) l2norm_bwd_kernel(const float* __restrict__ y, const float* __restrict__ inv_norm,
                                                         const float* __restrict__ dy, float* __restrict__ dx, int rows,
                                                         int cols) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float s = 0.0f;
  for (int c = lane; c < cols; c += 32) s += y[(int64_t)row * cols + c] * dy[(int64_t)row * cols + c];
  s = warp_sum(s);
  const float inv = inv_norm[row];
  for (int c = lane; c < cols; c += 32)
    dx[(int64_t)row * cols + c] = (dy[(int64_t)row * cols + c] - y[(int64_t)row * cols + c] * s) * inv;
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_embed_ln_scatter_fwd(const int64_t* ids, const int64_t* seg, const float* word, const float* pos,
                               const float* typ, const float* gamma, const float* beta, const float* vis, void* h,
                               float* mean, float* rstd, int B, int T, int H, int nvis, int vocab, float eps,
                               float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(ids && seg && word && pos && typ && gamma && beta && h && mean && rstd, "embed_fwd: null pointer");
  MMVQA_REQUIRE(B > 0 && T > 0 && H > 0 && nvis >= 0 && nvis <= T, "embed_fwd: bad shape");
  MMVQA_REQUIRE(nvis == 0 || vis, "embed_fwd: nvis > 0 needs vis");
  MMVQA_REQUIRE(dropout_p >= 0.0f && dropout_p < 1.0f, "embed_fwd: dropout_p must be in [0,1)");
  (void)vocab;
  const int64_t rows = (int64_t)B * T;
  const int grid = (int)((rows + 3) / 4);
  if (dtype == MMVQA_F32)
    embed_fwd_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(ids, seg, word, pos, typ, gamma, beta, vis, (float*)h, mean, rstd, B, T, H, nvis, eps, dropout_p, dropout_seed, g_seed_ctr);
  else if (dtype == MMVQA_BF16)
    embed_fwd_kernel<__nv_bfloat16><<<grid, 128, 0, as_stream(stream)>>>(ids, seg, word, pos, typ, gamma, beta, vis, (__nv_bfloat16*)h, mean, rstd, B, T, H, nvis, eps, dropout_p, dropout_seed, g_seed_ctr);
  else
    return set_err(MMVQA_ERR_ARG, "embed_fwd: bad dtype %d", dtype);
  MMVQA_LAUNCHED("embed_ln_scatter_fwd");
  return MMVQA_OK;
}

int mmvqa_embed_ln_scatter_bwd(const void* dh, const int64_t* ids, const int64_t* seg, const float* word,
                               const float* pos, const float* typ, const float* gamma, const float* mean,
                               const float* rstd, float* dword, float* dpos, float* dtyp, float* dgamma,
                               float* dbeta, float* dvis, int B, int T, int H, int nvis, int padding_idx,
                               float dropout_p, uint64_t dropout_seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(dh && ids && seg && word && pos && typ && gamma && mean && rstd, "embed_bwd: null pointer");
  MMVQA_REQUIRE(B > 0 && T > 0 && H > 0 && nvis >= 0 && nvis <= T, "embed_bwd: bad shape");
  const int64_t rows = (int64_t)B * T;
  int64_t want = (rows + 3) / 4, cap = (int64_t)num_sms() * 4;
  const int grid = (int)(want < cap ? want : cap);
  const size_t smem = sizeof(float) * 2 * (size_t)H;
  MMVQA_REQUIRE(smem <= 48 * 1024, "embed_bwd: H %d too large", H);
  auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (dtype == MMVQA_BF16 && H % 256 == 0 && H <= 1024 && al16(dh) && al16(word) && al16(pos) && al16(typ) && al16(gamma) &&
      al16(dword) && al16(dpos) && al16(dtyp) && al16(dgamma) && al16(dbeta) && al16(dvis)) {
    // 8-warp CTAs, two rows per warp: few enough CTAs that the per-CTA flush of the column sums stays cheap
    int64_t want2 = (rows + 15) / 16, cap2 = (int64_t)num_sms() * 2;
    const int grid2 = (int)(want2 < cap2 ? want2 : cap2);
    using B16 = __nv_bfloat16;
#define EMBED_BWD_VEC(KU) embed_bwd_vec_kernel<KU><<<grid2, 256, 0, as_stream(stream)>>>((const B16*)dh, ids, seg, word, pos, typ, gamma, mean, rstd, dword, dpos, dtyp, dgamma, dbeta, dvis, B, T, nvis, padding_idx, dropout_p, dropout_seed, g_seed_ctr)
    if (H == 256) EMBED_BWD_VEC(1); else if (H == 512) EMBED_BWD_VEC(2); else if (H == 768) EMBED_BWD_VEC(3); else EMBED_BWD_VEC(4);
#undef EMBED_BWD_VEC
    MMVQA_LAUNCHED("embed_ln_scatter_bwd");
    return MMVQA_OK;
  }
  if (dtype == MMVQA_F32)
    embed_bwd_kernel<float><<<grid, 128, smem, as_stream(stream)>>>((const float*)dh, ids, seg, word, pos, typ, gamma, mean, rstd, dword, dpos, dtyp, dgamma, dbeta, dvis, B, T, H, nvis, padding_idx, dropout_p, dropout_seed, g_seed_ctr);
  else if (dtype == MMVQA_BF16)
    embed_bwd_kernel<__nv_bfloat16><<<grid, 128, smem, as_stream(stream)>>>((const __nv_bfloat16*)dh, ids, seg, word, pos, typ, gamma, mean, rstd, dword, dpos, dtyp, dgamma, dbeta, dvis, B, T, H, nvis, padding_idx, dropout_p, dropout_seed, g_seed_ctr);
  else
    return set_err(MMVQA_ERR_ARG, "embed_bwd: bad dtype %d", dtype);
  MMVQA_LAUNCHED("embed_ln_scatter_bwd");
  return MMVQA_OK;
}

int mmvqa_masked_mean_fwd(const void* h, const float* mask, void* out, int B, int T, int H, int dtype,
                          mmvqa_stream_t stream) {
  MMVQA_REQUIRE(h && mask && out && B > 0 && T > 0 && H > 0, "masked_mean_fwd: bad args");
  dim3 grid((H + 255) / 256, B);
  if (dtype == MMVQA_F32)
    masked_mean_fwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)h, mask, (float*)out, T, H);
  else if (dtype == MMVQA_BF16)
    masked_mean_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)h, mask, (__nv_bfloat16*)out, T, H);
  else
    return set_err(MMVQA_ERR_ARG, "masked_mean_fwd: bad dtype %d", dtype);
  MMVQA_LAUNCHED("masked_mean_fwd");
  return MMVQA_OK;
}

int mmvqa_masked_mean_bwd(const void* dout, const float* mask, void* dh, int B, int T, int H, int dtype,
                          mmvqa_stream_t stream) {
  MMVQA_REQUIRE(dout && mask && dh && B > 0 && T > 0 && H > 0, "masked_mean_bwd: bad args");
  dim3 grid((H + 255) / 256, B);
  if (dtype == MMVQA_F32)
    masked_mean_bwd_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)dout, mask, (float*)dh, T, H);
  else if (dtype == MMVQA_BF16)
    masked_mean_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)dout, mask, (__nv_bfloat16*)dh, T, H);
  else
    return set_err(MMVQA_ERR_ARG, "masked_mean_bwd: bad dtype %d", dtype);
  MMVQA_LAUNCHED("masked_mean_bwd");
  return MMVQA_OK;
}

int mmvqa_l2norm_fwd(const float* x, float* y, float* inv_norm, int rows, int cols, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(x && y && rows >= 0 && cols > 0, "l2norm_fwd: bad args");
  if (rows == 0) return MMVQA_OK;
  l2norm_fwd_kernel<<<(rows + 3) / 4, 128, 0, as_stream(stream)>>>(x, y, inv_norm, rows, cols);
  MMVQA_LAUNCHED("l2norm_fwd");
  return MMVQA_OK;
}

int mmvqa_l2norm_bwd(const float* y, const float* inv_norm, const float* dy, float* dx, int rows, int cols,
                     mmvqa_stream_t stream) {
  MMVQA_REQUIRE(y && inv_norm && dy && dx && rows >= 0 && cols > 0, "l2norm_bwd: bad args");
  if (rows == 0) return MMVQA_OK;
  l2norm_bwd_kernel<<<(rows + 3) / 4, 128, 0, as_stream(stream)>>>(y, inv_norm, dy, dx, rows, cols);
  MMVQA_LAUNCHED("l2norm_bwd");
  return MMVQA_OK;
}

}  // extern "C"
