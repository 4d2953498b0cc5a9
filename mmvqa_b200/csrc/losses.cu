// losses.cu -- fused loss kernels (forward value and gradient in one launch):
//   * ASLSingleLabel           models/asl_singlelabel.py:23-52
//   * MLM cross entropy        pretrain/roco_utils.py:235-236  (log_softmax + NLL, every position)
//   * SupCon row pass          models/SupConLoss/loss.py:72-96
//   * multi-tensor Adam        vqamed2019/train.py:160 (torch.optim.Adam semantics)
#include "common.cuh"

namespace mmvqa {

// ---------------------------------------------------------------------------------
// ASL.  One block per sample.  With lp = log_softmax(x), p = exp(lp), y the target:
//   w_y = (1-p_y)^gp, w_c = p_c^gn (c != y);  ts_c = onehot*(1-eps) + eps/C
//   loss = -sum_c ts_c lp_c w_c
//   dloss/dx_k = -a_k + p_k * sum_c a_c,   a_c = ts_c w_c (1 + gn lp_c)            (c != y)
//                                           a_y = ts_y (w_y - gp lp_y p_y (1-p_y)^(gp-1))
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) asl_kernel(const T* __restrict__ logits, int64_t ld,
                                                  const int64_t* __restrict__ target, float* __restrict__ loss_rows,
                                                  float* __restrict__ dlogits, float* __restrict__ targets_classes, int C,
                                                  float gp, float gn, float eps) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const T* x = logits + (int64_t)b * ld;
  const int y = (int)target[b];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, to_f(x[c]));
  mx = block_max(mx, red);
  float se = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(to_f(x[c]) - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  const float smooth = eps > 0.0f ? eps / (float)C : 0.0f;
  const float keep = eps > 0.0f ? 1.0f - eps : 1.0f;
  float lsum = 0.0f, asum = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float lp = to_f(x[c]) - lse, p = expf(lp);
    const float ts = (c == y ? keep : 0.0f) + smooth;
    float w, a;
    if (c == y) {
      w = powf(1.0f - p, gp);
      a = ts * (gp == 0.0f ? w : w - gp * lp * p * powf(1.0f - p, gp - 1.0f));
    } else {
      w = gn == 0.0f ? 1.0f : expf(gn * lp);
      a = ts * w * (1.0f + gn * lp);
    }
    lsum += ts * lp * w;
    asum += a;
    if (targets_classes) targets_classes[(int64_t)b * C + c] = ts;
  }
  lsum = block_sum(lsum, red);
  asum = block_sum(asum, red);
  if (threadIdx.x == 0) loss_rows[b] = -lsum;
  if (dlogits) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float lp = to_f(x[c]) - lse, p = expf(lp);
      const float ts = (c == y ? keep : 0.0f) + smooth;
      float a;
      if (c == y) {
        const float w = powf(1.0f - p, gp);
        a = ts * (gp == 0.0f ? w : w - gp * lp * p * powf(1.0f - p, gp - 1.0f));
      } else {
        const float w = gn == 0.0f ? 1.0f : expf(gn * lp);
        a = ts * w * (1.0f + gn * lp);
      }
      dlogits[(int64_t)b * C + c] = -a + p * asum;
    }
  }
}

// ---------------------------------------------------------------------------------
// cross entropy per row: loss = lse - x_y;  dlogits = (softmax - onehot) * scale
// ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ce_kernel(const T* __restrict__ logits, int64_t ld,
                                                 const int64_t* __restrict__ target, float* __restrict__ loss_rows,
                                                 T* __restrict__ dlogits, int64_t ld_d, int C, float scale) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const T* x = logits + r * ld;
  const int y = (int)target[r];
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, to_f(x[c]));
  mx = block_max(mx, red);
  float se = 0.0f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += expf(to_f(x[c]) - mx);
  se = block_sum(se, red);
  const float lse = mx + logf(se);
  if (threadIdx.x == 0 && loss_rows) loss_rows[r] = lse - to_f(x[y]);
  if (dlogits) {
    T* d = dlogits + r * ld_d;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float p = expf(to_f(x[c]) - lse);
      d[c] = from_f<T>((p - (c == y ? 1.0f : 0.0f)) * scale);
    }
  }
}

// ---------------------------------------------------------------------------------
// cross entropy over a vocabulary that is never materialised: the [M, V] logits are produced in column chunks
// (vocab GEMM per chunk), pass 1 keeps a running (max, sum-exp, target logit) per row, pass 2 recomputes each chunk and
// turns it into the gradient chunk (models/mmbert.py:154-155 + pretrain/roco_utils.py:235-236)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_chunk_stats_kernel(const float* __restrict__ logits, int64_t ld,
                                                             const int64_t* __restrict__ target, int col0, int Vc,
                                                             float* __restrict__ rowmax, float* __restrict__ rowsum,
                                                             float* __restrict__ tgt_logit, int first) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const float* x = logits + r * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < Vc; c += blockDim.x) mx = fmaxf(mx, x[c]);
  mx = block_max(mx, red);
  float se = 0.0f;
  for (int c = threadIdx.x; c < Vc; c += blockDim.x) se += expf(x[c] - mx);
  se = block_sum(se, red);
  if (threadIdx.x == 0) {
    if (first) {
      rowmax[r] = mx;
      rowsum[r] = se;
    } else {
      const float m0 = rowmax[r], s0 = rowsum[r];
      const float m1 = fmaxf(m0, mx);
      rowmax[r] = m1;
      rowsum[r] = s0 * expf(m0 - m1) + se * expf(mx - m1);
    }
    const int y = (int)target[r] - col0;
    if (y >= 0 && y < Vc) tgt_logit[r] = x[y];
  }
}
template <typename O>
__global__ void __launch_bounds__(256) ce_chunk_grad_kernel(const float* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ target, int col0, int Vc,
                                                            const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                                            const float* __restrict__ row_scale, O* __restrict__ dl,
                                                            int64_t ld_d) {
  const int64_t r = blockIdx.x;
  const float* x = logits + r * ld;
  O* d = dl + r * ld_d;
  const float lse = rowmax[r] + logf(rowsum[r]);
  const float sc = row_scale[r];
  const int y = (int)target[r] - col0;
  for (int c = threadIdx.x; c < Vc; c += blockDim.x) d[c] = from_f<O>((expf(x[c] - lse) - (c == y ? 1.0f : 0.0f)) * sc);
}

// ---------------------------------------------------------------------------------
// SupCon: one block per anchor row of raw = anchor . contrast^T (not yet divided by temperature)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) supcon_rows_kernel(const float* __restrict__ raw, const float* __restrict__ mask,
                                                          float* __restrict__ loss_rows, float* __restrict__ G, int N,
                                                          int bsz, int row_offset, float temperature,
                                                          float base_temperature) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const int gi = row_offset + r;                 // global anchor index = its own column in the contrast set
  const float* x = raw + (int64_t)r * N;
  const float* mrow = mask ? mask + (int64_t)(gi % bsz) * bsz : nullptr;
  const float inv_t = 1.0f / temperature;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) mx = fmaxf(mx, x[j] * inv_t);   // max includes the self column
  mx = block_max(mx, red);
  float den = 0.0f, msum = 0.0f, mlog = 0.0f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    if (j == gi) continue;
    const float l = x[j] * inv_t - mx;
    den += expf(l);
    const float m = mrow ? mrow[j % bsz] : ((j % bsz) == (gi % bsz) ? 1.0f : 0.0f);
    msum += m;
    mlog += m * l;
  }
  den = block_sum(den, red);
  msum = block_sum(msum, red);
  mlog = block_sum(mlog, red);
  const float logden = logf(den);
  const float coef = -(temperature / base_temperature);
  if (threadIdx.x == 0) loss_rows[r] = coef * ((mlog - msum * logden) / msum);
  if (G) {
    float* g = G + (int64_t)r * N;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
      float v = 0.0f;
      if (j != gi) {
        const float l = x[j] * inv_t - mx;
        const float m = mrow ? mrow[j % bsz] : ((j % bsz) == (gi % bsz) ? 1.0f : 0.0f);
        v = coef * (m / msum - expf(l) / den) * inv_t;
      }
      g[j] = v;
    }
  }
}

// ---------------------------------------------------------------------------------
// multi-tensor Adam
// ---------------------------------------------------------------------------------
// one block per table entry (a chunk of one parameter tensor, built once on the host)
// (explicitly rounded operations: the dense and the row-gated loop must produce the same bits whatever the compiler's
// FMA contraction choices in either context)
__device__ __forceinline__ void adam_one(float& p, float& m, float& v, float g, float beta1, float beta2, float eps,
                                         float wd, float step, float bc2_sqrt, float grad_scale) {
  g = __fmul_rn(g, grad_scale);
  if (wd != 0.0f) g = __fmaf_rn(wd, p, g);
  m = __fmaf_rn(beta1, m, __fmul_rn(1.0f - beta1, g));
  v = __fmaf_rn(beta2, v, __fmul_rn(__fmul_rn(1.0f - beta2, g), g));
  p = __fsub_rn(p, __fdiv_rn(__fmul_rn(step, m), __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2_sqrt), eps)));
}

// U = independent 16-byte groups per thread and iteration.  U = 1 ("background"): a gentle stream of few bytes in flight per
// CTA for the per-layer updates that run underneath the latency-bound backward pass (the U = 2 loop finishes each update
// sooner but slowed the whole step by ~0.1 ms).  U = 2: updates on the critical path (the tail of the step, plain step()).
template <int U>
__global__ void __launch_bounds__(256, 3) adam_kernel(const mmvqa_adam_desc* __restrict__ table, float lr, float beta1,
                                                   float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                                                   float grad_scale, const int* __restrict__ step_dev, int n_chunks,
                                                   const float* __restrict__ hyper_dev) {
  pdl_wait();
  // (no early launch_dependents: this kernel rewrites the bf16 weight copies that a dependent GEMM launched with
  // b_static = 1 would prefetch before its own griddepcontrol.wait; the implicit trigger at completion is the safe one)
  if (hyper_dev) {   // lr / grad_scale refreshed by the host between graph replays
    lr = hyper_dev[0];
    grad_scale = hyper_dev[1];
  }
  if (step_dev) {  // step counter lives on the device (CUDA-graph replay safe)
    const float t = (float)(*step_dev);
    bc1 = 1.0f - powf(beta1, t);
    bc2_sqrt = sqrtf(1.0f - powf(beta2, t));
  }
  const float step = lr / bc1;
  // grid-stride over the chunk table: a capped grid bounds the HBM bandwidth this launch takes while it runs
  // concurrently with the latency-bound backward kernels
  for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
  const mmvqa_adam_desc d = table[ch];
  const bool g16 = (d.flags & 1) != 0;
  if (d.row_live != nullptr) {
    // row-gated chunk (embedding table): rows that never received gradient are skipped (identity update)
    const int nrows = (int)(d.n / d.row_len), r4 = (int)(d.row_len / 4);
    // the CTA first compacts the live rows of the chunk (in blocks of 1024 flags), then ALL its threads share their
    // 16-byte groups: a warp walking its own rows paid ~6 dependent HBM round trips per row (20 us for a table update
    // that moves 1 MB, at the very end of the step)
    __shared__ int live_rows[1024];
    __shared__ int n_live;
    for (int rb = 0; rb < nrows; rb += 1024) {
      __syncthreads();
      if (threadIdx.x == 0) n_live = 0;
      __syncthreads();
      for (int r = rb + threadIdx.x; r < min(nrows, rb + 1024); r += blockDim.x)
        if (d.row_live[r] != 0) live_rows[atomicAdd(&n_live, 1)] = r;
      __syncthreads();
      const int items = n_live * r4;
      for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int li = it / r4;
        const int64_t q = (int64_t)live_rows[li] * r4 + (it - li * r4);
        float4 p = __ldcs(reinterpret_cast<const float4*>(d.p) + q), m = __ldcs(reinterpret_cast<const float4*>(d.m) + q);
        float4 v = __ldcs(reinterpret_cast<const float4*>(d.v) + q);
        float4 g;
        if (g16) {
          const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(d.g) + q);
          g.x = __uint_as_float(raw.x << 16); g.y = __uint_as_float(raw.x & 0xffff0000u);
          g.z = __uint_as_float(raw.y << 16); g.w = __uint_as_float(raw.y & 0xffff0000u);
        } else {
          g = __ldcs(reinterpret_cast<const float4*>(d.g) + q);
        }
        adam_one(p.x, m.x, v.x, g.x, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
        adam_one(p.y, m.y, v.y, g.y, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
        adam_one(p.z, m.z, v.z, g.z, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
        adam_one(p.w, m.w, v.w, g.w, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
        __stcs(reinterpret_cast<float4*>(d.p) + q, p);
        __stcs(reinterpret_cast<float4*>(d.m) + q, m);
        __stcs(reinterpret_cast<float4*>(d.v) + q, v);
        if (d.bf16_out) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
          uint2 o;
          o.x = *reinterpret_cast<uint32_t*>(&lo);
          o.y = *reinterpret_cast<uint32_t*>(&hi);
          __stcs(reinterpret_cast<uint2*>(d.bf16_out) + q, o);
        }
      }
    }
    continue;
  }
  const float* g32 = reinterpret_cast<const float*>(d.g);
  const __nv_bfloat16* gb = reinterpret_cast<const __nv_bfloat16*>(d.g);
  const bool vec = ((reinterpret_cast<uintptr_t>(d.p) | reinterpret_cast<uintptr_t>(d.m) | reinterpret_cast<uintptr_t>(d.v)) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(d.g) & (g16 ? 7 : 15)) == 0 &&
                   (d.bf16_out == nullptr || (reinterpret_cast<uintptr_t>(d.bf16_out) & 7) == 0);
  int64_t done = 0;
  if (vec) {
    const int64_t n4 = d.n / 4;
    // U independent 16-byte groups per thread and iteration: all loads are issued before the first use, so a CTA
    // pays ~one memory latency per 2048 elements (at <= 80 registers: its CTAs stay small tenants of an SM) (a one-group loop made every chunk a chain of 32 dependent HBM round
    // trips: ~35 us per launch however small, which is what the updates at the tail of the step cost)
    for (int64_t i0 = threadIdx.x; i0 < n4; i0 += (int64_t)blockDim.x * U) {
      float4 p[U], m[U], v[U], g[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * blockDim.x;
        if (i < n4) {
          // evict-first (streaming) accesses: 28 bytes per parameter pass through exactly once; keeping them out of the
          // L2 working set matters when the update runs underneath the backward pass, whose saved activations live there
          p[u] = __ldcs(reinterpret_cast<const float4*>(d.p) + i);
          m[u] = __ldcs(reinterpret_cast<const float4*>(d.m) + i);
          v[u] = __ldcs(reinterpret_cast<const float4*>(d.v) + i);
          if (g16) {
            const uint2 raw = __ldcs(reinterpret_cast<const uint2*>(gb) + i);
            g[u].x = __uint_as_float(raw.x << 16); g[u].y = __uint_as_float(raw.x & 0xffff0000u);
            g[u].z = __uint_as_float(raw.y << 16); g[u].w = __uint_as_float(raw.y & 0xffff0000u);
          } else {
            g[u] = __ldcs(reinterpret_cast<const float4*>(g32) + i);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * blockDim.x;
        if (i < n4) {
          adam_one(p[u].x, m[u].x, v[u].x, g[u].x, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
          adam_one(p[u].y, m[u].y, v[u].y, g[u].y, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
          adam_one(p[u].z, m[u].z, v[u].z, g[u].z, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
          adam_one(p[u].w, m[u].w, v[u].w, g[u].w, beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
          __stcs(reinterpret_cast<float4*>(d.p) + i, p[u]);
          __stcs(reinterpret_cast<float4*>(d.m) + i, m[u]);
          __stcs(reinterpret_cast<float4*>(d.v) + i, v[u]);
          if (d.bf16_out) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p[u].x, p[u].y), hi = __floats2bfloat162_rn(p[u].z, p[u].w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo);
            o.y = *reinterpret_cast<uint32_t*>(&hi);
            __stcs(reinterpret_cast<uint2*>(d.bf16_out) + i, o);
          }
        }
      }
    }
    done = n4 * 4;
  }
  for (int64_t i = done + threadIdx.x; i < d.n; i += blockDim.x) {
    float p = d.p[i], m = d.m[i], v = d.v[i];
    adam_one(p, m, v, g16 ? __bfloat162float(gb[i]) : g32[i], beta1, beta2, eps, wd, step, bc2_sqrt, grad_scale);
    d.p[i] = p;
    d.m[i] = m;
    d.v[i] = v;
    if (d.bf16_out) reinterpret_cast<__nv_bfloat16*>(d.bf16_out)[i] = __float2bfloat16_rn(p);
  }
  }
}

__global__ void __launch_bounds__(256) mark_rows_kernel(unsigned char* __restrict__ row_live, const int64_t* __restrict__ ids,
                                                        int64_t n, int64_t rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t id = ids[i];
  if (id >= 0 && id < rows) row_live[id] = 1;
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_mark_rows(unsigned char* row_live, const int64_t* ids, int64_t n, int64_t rows, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(row_live && (ids || n == 0) && n >= 0 && rows > 0, "mark_rows: bad args");
  if (n == 0) return MMVQA_OK;
  mark_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(row_live, ids, n, rows);
  MMVQA_LAUNCHED("mark_rows");
  return MMVQA_OK;
}

int mmvqa_asl_fwd_bwd(const void* logits, int64_t ld, const int64_t* target, float* loss_rows, float* dlogits,
                      float* targets_classes, int B, int C, float gamma_pos, float gamma_neg, float eps, int dtype,
                      mmvqa_stream_t stream) {
  MMVQA_REQUIRE(logits && target && loss_rows && B >= 0 && C > 0 && ld >= C, "asl: bad args");
  if (B == 0) return MMVQA_OK;
  if (dtype == MMVQA_F32)
    asl_kernel<float><<<B, 256, 0, as_stream(stream)>>>((const float*)logits, ld, target, loss_rows, dlogits, targets_classes, C, gamma_pos, gamma_neg, eps);
  else if (dtype == MMVQA_BF16)
    asl_kernel<__nv_bfloat16><<<B, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)logits, ld, target, loss_rows, dlogits, targets_classes, C, gamma_pos, gamma_neg, eps);
  else
    return set_err(MMVQA_ERR_ARG, "asl: bad dtype %d", dtype);
  MMVQA_LAUNCHED("asl_fwd_bwd");
  return MMVQA_OK;
}

int mmvqa_ce_fwd_bwd(const void* logits, int64_t ld, const int64_t* target, float* loss_rows, void* dlogits,
                     int64_t ld_d, int64_t rows, int C, float scale, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(logits && target && rows >= 0 && C > 0 && ld >= C, "ce: bad args");
  MMVQA_REQUIRE(!dlogits || ld_d >= C, "ce: bad ld_d");
  MMVQA_REQUIRE(rows <= 2147483647LL, "ce: too many rows");
  if (rows == 0) return MMVQA_OK;
  if (dtype == MMVQA_F32)
    ce_kernel<float><<<(int)rows, 256, 0, as_stream(stream)>>>((const float*)logits, ld, target, loss_rows, (float*)dlogits, ld_d, C, scale);
  else if (dtype == MMVQA_BF16)
    ce_kernel<__nv_bfloat16><<<(int)rows, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)logits, ld, target, loss_rows, (__nv_bfloat16*)dlogits, ld_d, C, scale);
  else
    return set_err(MMVQA_ERR_ARG, "ce: bad dtype %d", dtype);
  MMVQA_LAUNCHED("ce_fwd_bwd");
  return MMVQA_OK;
}

int mmvqa_ce_chunk_stats(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int col0, int Vc,
                         float* rowmax, float* rowsum, float* tgt_logit, int first, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(logits && target && rowmax && rowsum && tgt_logit && rows >= 0 && Vc > 0 && ld >= Vc && col0 >= 0,
                "ce_chunk_stats: bad args");
  MMVQA_REQUIRE(rows <= 2147483647LL, "ce_chunk_stats: too many rows");
  if (rows == 0) return MMVQA_OK;
  ce_chunk_stats_kernel<<<(int)rows, 256, 0, as_stream(stream)>>>(logits, ld, target, col0, Vc, rowmax, rowsum, tgt_logit, first);
  MMVQA_LAUNCHED("ce_chunk_stats");
  return MMVQA_OK;
}

int mmvqa_ce_chunk_grad(const float* logits, int64_t ld, const int64_t* target, int64_t rows, int col0, int Vc,
                        const float* rowmax, const float* rowsum, const float* row_scale, void* dlogits, int64_t ld_d,
                        int dl_dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(logits && target && rowmax && rowsum && row_scale && dlogits && rows >= 0 && Vc > 0 && ld >= Vc && ld_d >= Vc,
                "ce_chunk_grad: bad args");
  MMVQA_REQUIRE(rows <= 2147483647LL, "ce_chunk_grad: too many rows");
  if (rows == 0) return MMVQA_OK;
  if (dl_dtype == MMVQA_F32)
    ce_chunk_grad_kernel<float><<<(int)rows, 256, 0, as_stream(stream)>>>(logits, ld, target, col0, Vc, rowmax, rowsum, row_scale, (float*)dlogits, ld_d);
  else if (dl_dtype == MMVQA_BF16)
    ce_chunk_grad_kernel<__nv_bfloat16><<<(int)rows, 256, 0, as_stream(stream)>>>(logits, ld, target, col0, Vc, rowmax, rowsum, row_scale, (__nv_bfloat16*)dlogits, ld_d);
  else
    return set_err(MMVQA_ERR_ARG, "ce_chunk_grad: bad dtype %d", dl_dtype);
  MMVQA_LAUNCHED("ce_chunk_grad");
  return MMVQA_OK;
}

int mmvqa_supcon_rows(const float* logits, const float* mask, float* loss_rows, float* G, int R, int N, int bsz,
                      int row_offset, float temperature, float base_temperature, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(logits && loss_rows && R >= 0 && N > 0 && bsz > 0 && N % bsz == 0, "supcon_rows: bad args");
  MMVQA_REQUIRE(row_offset >= 0 && row_offset + R <= N, "supcon_rows: anchor rows out of range");
  if (R == 0) return MMVQA_OK;
  supcon_rows_kernel<<<R, 256, 0, as_stream(stream)>>>(logits, mask, loss_rows, G, N, bsz, row_offset, temperature, base_temperature);
  MMVQA_LAUNCHED("supcon_rows");
  return MMVQA_OK;
}

int mmvqa_adam_step(const mmvqa_adam_desc* table, int n_chunks, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int step, const int* step_dev, float grad_scale, int max_ctas, int background,
                    mmvqa_stream_t stream) {
  MMVQA_REQUIRE(table && n_chunks >= 0 && (step >= 1 || step_dev), "adam: bad args");
  if (step < 1) step = 1;
  if (n_chunks == 0) return MMVQA_OK;
  const float bc1 = 1.0f - powf(beta1, (float)step);
  const float bc2 = 1.0f - powf(beta2, (float)step);
  const int grid = (max_ctas > 0 && max_ctas < n_chunks) ? max_ctas : n_chunks;
  if (background)
    MMVQA_CUDA(launch_pdl(adam_kernel<1>, dim3(grid), dim3(256), 0, as_stream(stream), table, lr, beta1, beta2, eps, weight_decay,
                          bc1, sqrtf(bc2), grad_scale, step_dev, n_chunks, (const float*)nullptr));
  else
    MMVQA_CUDA(launch_pdl(adam_kernel<2>, dim3(grid), dim3(256), 0, as_stream(stream), table, lr, beta1, beta2, eps, weight_decay,
                          bc1, sqrtf(bc2), grad_scale, step_dev, n_chunks, (const float*)nullptr));
  MMVQA_LAUNCHED("adam_step");
  return MMVQA_OK;
}

int mmvqa_adam_step_dev(const mmvqa_adam_desc* table, int n_chunks, const float* hyper_dev, float beta1, float beta2,
                        float eps, float weight_decay, const int* step_dev, int max_ctas, int background,
                        mmvqa_stream_t stream) {
  MMVQA_REQUIRE(table && n_chunks >= 0 && hyper_dev && step_dev, "adam_dev: bad args");
  if (n_chunks == 0) return MMVQA_OK;
  const int grid = (max_ctas > 0 && max_ctas < n_chunks) ? max_ctas : n_chunks;
  if (background)
    MMVQA_CUDA(launch_pdl(adam_kernel<1>, dim3(grid), dim3(256), 0, as_stream(stream), table, 0.0f, beta1, beta2, eps, weight_decay,
                          1.0f, 1.0f, 1.0f, step_dev, n_chunks, hyper_dev));
  else
    MMVQA_CUDA(launch_pdl(adam_kernel<2>, dim3(grid), dim3(256), 0, as_stream(stream), table, 0.0f, beta1, beta2, eps, weight_decay,
                          1.0f, 1.0f, 1.0f, step_dev, n_chunks, hyper_dev));
  MMVQA_LAUNCHED("adam_step");
  return MMVQA_OK;
}

}  // extern "C"
