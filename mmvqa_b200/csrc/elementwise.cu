// elementwise.cu -- HBM-bound kernels of the fusion-encoder path: bias+activation, column sums,
// casts, residual+LayerNorm (fwd/bwd), dropout.  All use 128-bit vectorised global accesses when
// the row length and pointers allow it and fall back to scalar accesses otherwise.
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace mmvqa {

thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};
const unsigned long long* g_seed_ctr = nullptr;

int set_err(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

bool pdl_enabled() {
  // Programmatic dependent launch (griddepcontrol): every kernel of the library is launched so that it may become
  // resident while its predecessor on the stream is still running; set-up (barriers, TMEM allocation, tensor-map
  // and weight-tile prefetch) overlaps the predecessor, dependent memory is touched only after griddepcontrol.wait.
  // Measured on the captured B=16 step: 3.31 -> 3.24 ms once the GEMM rings are capped so two CTAs share an SM.
  // MMVQA_NO_PDL=1 turns it off (A/B runs).
  static const bool on = getenv("MMVQA_NO_PDL") == nullptr;
  return on;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------
// bias + activation
// ---------------------------------------------------------------------------------
template <typename T, bool BWD, bool VEC>
__global__ void __launch_bounds__(256) bias_act_kernel(const T* __restrict__ x, const float* __restrict__ bias,
                                                       const T* __restrict__ dy, T* __restrict__ out, int64_t rows,
                                                       int cols, int act) {
  if (VEC) {
    constexpr int N = Vec16<T>::N;
    const int vcols = cols / N;
    const int64_t total = rows * (int64_t)vcols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      int c0 = (int)(i % vcols) * N;
      Vec16<T> v, g, o;
      v.load(x + i * N);
      if (BWD) g.load(dy + i * N);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float z = v.get(j) + (bias ? __ldg(bias + c0 + j) : 0.0f);
        o.set(j, BWD ? g.get(j) * dact_rt(act, z) : act_rt(act, z));
      }
      o.store(out + i * N);
    }
  } else {
    const int64_t total = rows * (int64_t)cols;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      int c = (int)(i % cols);
      float z = to_f(x[i]) + (bias ? __ldg(bias + c) : 0.0f);
      out[i] = from_f<T>(BWD ? to_f(dy[i]) * dact_rt(act, z) : act_rt(act, z));
    }
  }
}

template <typename T, bool BWD>
static int bias_act_launch(const void* x, const float* bias, const void* dy, void* out, int64_t rows, int cols, int act,
                           cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return MMVQA_OK;
  bool vec = (cols % Vec16<T>::N == 0) && aligned16(x) && aligned16(out) && (!BWD || aligned16(dy));
  int64_t work = rows * (int64_t)cols / (vec ? Vec16<T>::N : 1);
  int grid = (int)((work + 255) / 256 < (int64_t)num_sms() * 16 ? (work + 255) / 256 : (int64_t)num_sms() * 16);
  if (vec)
    bias_act_kernel<T, BWD, true><<<grid, 256, 0, st>>>((const T*)x, bias, (const T*)dy, (T*)out, rows, cols, act);
  else
    bias_act_kernel<T, BWD, false><<<grid, 256, 0, st>>>((const T*)x, bias, (const T*)dy, (T*)out, rows, cols, act);
  MMVQA_LAUNCHED("bias_act");
  return MMVQA_OK;
}

// ---------------------------------------------------------------------------------
// column sums (bias gradients):  out[c] = sum_r x[r, c]
// block = 32 columns x 8 row lanes; grid.y splits the rows; partial sums land with one atomic per
// (block, column) into the zero-filled output.
// ---------------------------------------------------------------------------------
// L2 prefetch hints: every thread walks 128-byte lines of the listed ranges (grid-stride over all ranges at once)
__global__ void __launch_bounds__(256) l2_prefetch_kernel(const mmvqa_prefetch_list list) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  for (int r = 0; r < list.n; ++r) {
    const char* base = reinterpret_cast<const char*>(list.ptr[r]);
    const int64_t lines = (list.bytes[r] + 127) >> 7;
    for (int64_t i = tid; i < lines; i += nthr) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (i << 7)));
  }
}

template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int64_t ldx, float* __restrict__ out,
                                                     int64_t rows, int cols, int64_t rows_per_block) {
  __shared__ float red[8][33];
  int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  int c = blockIdx.x * 32 + tx;
  int64_t r0 = blockIdx.y * rows_per_block, r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s = 0.0f;
  if (c < cols)
    for (int64_t r = r0 + ty; r < r1; r += 8) s += to_f(x[r * ldx + c]);
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t += red[j][tx];
    if (gridDim.y == 1) out[c] = t;     // single row split: plain store, no zero-fill needed
    else atomicAdd(out + c, t);
  }
}

// ---------------------------------------------------------------------------------
// casts
// ---------------------------------------------------------------------------------
template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    if (i + 3 < n) {
      float a = to_f(src[i]), b = to_f(src[i + 1]), c = to_f(src[i + 2]), d = to_f(src[i + 3]);
      dst[i] = from_f<D>(a); dst[i + 1] = from_f<D>(b); dst[i + 2] = from_f<D>(c); dst[i + 3] = from_f<D>(d);
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = from_f<D>(to_f(src[j]));
    }
  }
}
// specialised hot case: contiguous fp32 -> bf16 with 128-bit loads, 64-bit stores
__global__ void __launch_bounds__(256) cast_f32_bf16_vec(const float4* __restrict__ src, uint2* __restrict__ dst,
                                                         int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = src[i];
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    dst[i] = o;
  }
}

template <typename S, typename D>
__global__ void __launch_bounds__(256) cast_pad_kernel(const S* __restrict__ src, int64_t ld_src, D* __restrict__ dst,
                                                       int64_t ld_dst, int64_t rows, int cols) {
  int64_t total = rows * ld_dst;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / ld_dst;
    int c = (int)(i - r * ld_dst);
    dst[i] = from_f<D>(c < cols ? to_f(src[r * ld_src + c]) : 0.0f);
  }
}

// multi-problem fp32 -> bf16 cast_pad: a thread converts 4 consecutive destination elements (one 8-byte store); the
// source is read with one 16-byte load when its rows allow it.  unit_end[i] = running total of 4-element units.
struct CastMultiArgs {
  mmvqa_cast_list l;
  int64_t unit_end[MMVQA_CAST_MULTI_MAX];
  int vec_src[MMVQA_CAST_MULTI_MAX];
};
__global__ void __launch_bounds__(256) cast_pad_multi_kernel(const __grid_constant__ CastMultiArgs a) {
  const int64_t total = a.unit_end[a.l.n - 1];
  for (int64_t u = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; u < total; u += (int64_t)gridDim.x * blockDim.x) {
    int s = 0;
    while (u >= a.unit_end[s]) ++s;
    const int64_t v = u - (s ? a.unit_end[s - 1] : 0);
    const int64_t upr = a.l.ld_dst[s] >> 2;          // units per destination row
    const int64_t r = v / upr;
    const int c = (int)(v - r * upr) << 2;
    const int cols = a.l.cols[s];
    if (a.l.src_bf16[s]) {      // bf16 rows re-padded to the TMA leading dimension: copy the bits
      const unsigned short* bp = reinterpret_cast<const unsigned short*>(a.l.src[s]) + r * a.l.ld_src[s] + c;
      uint2 o;
      if (a.vec_src[s] && c + 3 < cols) {
        o = *reinterpret_cast<const uint2*>(bp);
      } else {
        const uint32_t e0 = c < cols ? bp[0] : 0u, e1 = c + 1 < cols ? bp[1] : 0u, e2 = c + 2 < cols ? bp[2] : 0u,
                       e3 = c + 3 < cols ? bp[3] : 0u;
        o.x = e0 | (e1 << 16);
        o.y = e2 | (e3 << 16);
      }
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.l.dst[s]) + r * a.l.ld_dst[s] + c) = o;
      continue;
    }
    const float* sp = reinterpret_cast<const float*>(a.l.src[s]) + r * a.l.ld_src[s] + c;
    float4 x;
    if (a.vec_src[s] && c + 3 < cols) {
      x = __ldcs(reinterpret_cast<const float4*>(sp));
    } else {
      x.x = c < cols ? sp[0] : 0.0f;
      x.y = c + 1 < cols ? sp[1] : 0.0f;
      x.z = c + 2 < cols ? sp[2] : 0.0f;
      x.w = c + 3 < cols ? sp[3] : 0.0f;
    }
    __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&lo);
    o.y = *reinterpret_cast<uint32_t*>(&hi);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.l.dst[s]) + r * a.l.ld_dst[s] + c) = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) scale_kernel(T* __restrict__ x, const float* __restrict__ scalar, float host_factor,
                                                    int64_t n) {
  float s = (scalar ? *scalar : 1.0f) * host_factor;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = from_f<T>(to_f(x[i]) * s);
}

// y = x * keep / (1-p), keep from the counter hash of (seed, element index)
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t n, float p,
                                                      uint64_t seed, const unsigned long long* seed_ctr) {
  seed = seed_eff(seed, seed_ctr);
  uint32_t thr = (uint32_t)(p * 4294967296.0);
  float inv = 1.0f / (1.0f - p);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = from_f<T>(hash32(seed, (uint64_t)i) >= thr ? to_f(x[i]) * inv : 0.0f);
}

// ---------------------------------------------------------------------------------
// residual + LayerNorm.  One warp per row.  Fast path (cols <= 1024): the row lives in registers,
// every global access is a 128-bit vector (VEC) or a scalar (unaligned / odd row length).
// Generic path (cols > 1024): re-reads the row.  Two-pass mean / variance like ATen.
// ---------------------------------------------------------------------------------
constexpr int LN_CACHE = 32;  // values per lane -> rows up to 1024 columns stay in registers

template <typename T, bool VEC>
__global__ void __launch_bounds__(128) add_ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ y, T* __restrict__ sum_out,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         int64_t rows, int cols, float eps) {
  constexpr int N = VEC ? Vec16<T>::N : 1;
  constexpr int ITER = LN_CACHE / N;
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= rows) return;
  const T* xr = x + row * cols;
  const T* rr = res ? res + row * cols : nullptr;
  T* so = sum_out ? sum_out + row * cols : nullptr;
  float v[LN_CACHE];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
      if (VEC) {
        Vec16<T> a, b;
        a.load(xr + c);
        if (rr) b.load(rr + c);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          float t = a.get(j) + (rr ? b.get(j) : 0.0f);
          if (so) {  // normalise the value that is stored (bf16-rounded) so backward sees the same row
            a.set(j, t);
            t = a.get(j);
          }
          v[k * N + j] = t;
          s += t;
        }
        if (so) a.store(so + c);
      } else {
        float t = to_f(xr[c]) + (rr ? to_f(rr[c]) : 0.0f);
        if (so) {
          T q = from_f<T>(t);
          so[c] = q;
          t = to_f(q);
        }
        v[k] = t;
        s += t;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) v[k * N + j] = 0.0f;
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float d = v[k * N + j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  T* yr = y + row * cols;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
      if (VEC) {
        Vec16<T> o;
#pragma unroll
        for (int j = 0; j < N; ++j)
          o.set(j, (v[k * N + j] - mean) * rstd * __ldg(gamma + c + j) + __ldg(beta + c + j));
        o.store(yr + c);
      } else {
        yr[c] = from_f<T>((v[k] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
      }
    }
  }
}

// Vector path used whenever cols % (16 B) == 0 and cols <= 1024: the row stays in registers as the RAW 128-bit
// vectors (12 registers for a 768-wide bf16 row instead of 24-32 floats) and is converted on the fly in each of the
// three passes -- the ALU has slack, HBM does not.  Low register count -> 8 warps x 8 CTAs per SM in flight.
template <typename T>
__global__ void __launch_bounds__(256) add_ln_fwd_packed(const T* __restrict__ x, const T* __restrict__ res,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         T* __restrict__ y, T* __restrict__ sum_out,
                                                         float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                         int64_t rows, int cols, float eps) {
  constexpr int N = Vec16<T>::N;
  constexpr int ITER = LN_CACHE / N;
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= rows) return;
  const T* xr = x + row * cols;
  const T* rr = res ? res + row * cols : nullptr;
  T* so = sum_out ? sum_out + row * cols : nullptr;
  Vec16<T> raw[ITER];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) raw[k].load(xr + c);
  }
  if (rr) {
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < cols) {
        Vec16<T> b;
        b.load(rr + c);
#pragma unroll
        for (int j = 0; j < N; ++j) raw[k].set(j, raw[k].get(j) + b.get(j));   // the stored (rounded) sum is what is normalised
        if (so) raw[k].store(so + c);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
#pragma unroll
      for (int j = 0; j < N; ++j) s += raw[k].get(j);
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float d = raw[k].get(j) - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  T* yr = y + row * cols;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c + j));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c + j));
        o.set(j, (raw[k].get(j) - mean) * rstd * g4.x + b4.x);
        o.set(j + 1, (raw[k].get(j + 1) - mean) * rstd * g4.y + b4.y);
        o.set(j + 2, (raw[k].get(j + 2) - mean) * rstd * g4.z + b4.z);
        o.set(j + 3, (raw[k].get(j + 3) - mean) * rstd * g4.w + b4.w);
      }
      o.store(yr + c);
    }
  }
}

// LayerNorm over the sum of fp32 split-K partial tiles (mmvqa_gemm with c_split_stride): the reduction of the
// slabs, the dropout of the branch, the residual add and the normalisation are ONE pass, so a K = 3072 GEMM at
// M = 448 can spread over every SM without an atomic or a second sweep.  s = dropout(sum_i parts[i]) + res is
// rounded to T and stored (sum_out) -- it is what the backward pass normalises again.
// NP > 0: the slab count is a compile-time constant, so every slab load of a row is issued before the first use (the
// run-time loop made a row a chain of nparts x 3 dependent L2 round trips: 10 us for a kernel that moves 5 MB).
template <typename T, int NP>
__global__ void __launch_bounds__(256) add_ln_fwd_parts(const float* __restrict__ parts, int nparts, int64_t part_stride,
                                                        const T* __restrict__ res, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, T* __restrict__ y,
                                                        T* __restrict__ sum_out, float* __restrict__ mean_out,
                                                        float* __restrict__ rstd_out, int64_t rows, int cols, float eps,
                                                        float p, unsigned long long seed,
                                                        const unsigned long long* seed_ctr) {
  constexpr int N = Vec16<T>::N;
  constexpr int ITER = LN_CACHE / N;
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= rows) return;
  const bool drop = p > 0.0f;
  if (drop) seed = seed_eff(seed, seed_ctr);
  const uint32_t thr = (uint32_t)(p * 4294967296.0);
  const float inv_keep = drop ? 1.0f / (1.0f - p) : 1.0f;
  const T* rr = res ? res + row * cols : nullptr;
  T* so = sum_out ? sum_out + row * cols : nullptr;
  Vec16<T> raw[ITER];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
      float acc[N];
#pragma unroll
      for (int j = 0; j < N; ++j) acc[j] = 0.0f;
      if (NP > 0) {
        float4 t[NP > 0 ? NP : 1][N / 4];
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          const float* pr = parts + (int64_t)i * part_stride + row * cols + c;
#pragma unroll
          for (int j = 0; j < N / 4; ++j) t[i][j] = __ldg(reinterpret_cast<const float4*>(pr) + j);
        }
#pragma unroll
        for (int i = 0; i < NP; ++i) {      // same summation order as the run-time loop
#pragma unroll
          for (int j = 0; j < N / 4; ++j) {
            acc[4 * j] += t[i][j].x; acc[4 * j + 1] += t[i][j].y; acc[4 * j + 2] += t[i][j].z; acc[4 * j + 3] += t[i][j].w;
          }
        }
      } else {
        for (int i = 0; i < nparts; ++i) {
          const float* pr = parts + (int64_t)i * part_stride + row * cols + c;
#pragma unroll
          for (int j = 0; j < N; j += 4) {
            const float4 t = *reinterpret_cast<const float4*>(pr + j);
            acc[j] += t.x; acc[j + 1] += t.y; acc[j + 2] += t.z; acc[j + 3] += t.w;
          }
        }
      }
      Vec16<T> b;
      if (rr) b.load(rr + c);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float v = acc[j];
        if (drop) v = hash32(seed, (uint64_t)(row * cols + c + j)) >= thr ? v * inv_keep : 0.0f;
        raw[k].set(j, v + (rr ? b.get(j) : 0.0f));
      }
      if (so) raw[k].store(so + c);
#pragma unroll
      for (int j = 0; j < N; ++j) s += raw[k].get(j);
    }
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.0f;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float d = raw[k].get(j) - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  T* yr = y + row * cols;
#pragma unroll
  for (int k = 0; k < ITER; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c + j));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c + j));
        o.set(j, (raw[k].get(j) - mean) * rstd * g4.x + b4.x);
        o.set(j + 1, (raw[k].get(j + 1) - mean) * rstd * g4.y + b4.y);
        o.set(j + 2, (raw[k].get(j + 2) - mean) * rstd * g4.z + b4.z);
        o.set(j + 3, (raw[k].get(j + 3) - mean) * rstd * g4.w + b4.w);
      }
      o.store(yr + c);
    }
  }
}

// generic (any cols): re-reads the row from global memory
template <typename T>
__global__ void __launch_bounds__(128) add_ln_fwd_generic(const T* __restrict__ x, const T* __restrict__ res,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          T* __restrict__ y, T* __restrict__ sum_out,
                                                          float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                          int64_t rows, int cols, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  pdl_wait();
  pdl_trigger();
  if (row >= rows) return;
  const T* xr = x + row * cols;
  const T* rr = res ? res + row * cols : nullptr;
  T* so = sum_out ? sum_out + row * cols : nullptr;
  float s = 0.0f;
  for (int c = lane; c < cols; c += 32) {
    float t = to_f(xr[c]) + (rr ? to_f(rr[c]) : 0.0f);
    if (so) {
      T q = from_f<T>(t);
      so[c] = q;
      t = to_f(q);
    }
    s += t;
  }
  const float mean = warp_sum(s) / (float)cols;
  float q = 0.0f;
  for (int c = lane; c < cols; c += 32) {
    float t = so ? to_f(so[c]) : to_f(xr[c]) + (rr ? to_f(rr[c]) : 0.0f);
    q += (t - mean) * (t - mean);
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)cols + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  for (int c = lane; c < cols; c += 32) {
    float t = so ? to_f(so[c]) : to_f(xr[c]) + (rr ? to_f(rr[c]) : 0.0f);
    y[row * cols + c] = from_f<T>((t - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c));
  }
}

struct LnBwdExtra {
  void* dx_drop;            // optional: dropout(dx), same dtype
  float* dxsum;             // optional: += column sums of dx_drop (or of dx when dx_drop == NULL)
  float p;
  unsigned long long seed;
  const unsigned long long* seed_ctr;
  // optional (packed kernel only): the incoming gradient is sum_i dy_parts[i] (fp32 split-K slabs of the dgrad GEMM)
  // + dy (the residual branch, may be NULL), rounded to T -- exactly what a dgrad GEMM with a residual epilogue stores
  const float* dy_parts;
  int nparts;
  long long part_stride;
  // optional (packed kernel only): every CTA stores its [3][cols] column sums (dgamma | dbeta | dxsum) to
  // partials + blockIdx.x * 3 * cols with plain stores instead of atomics; mmvqa_ln_partials_reduce folds them later
  float* partials;
};

// Packed backward (vector path): the two input rows stay in registers as raw 128-bit vectors and are re-expanded in
// the second pass instead of keeping 2 x 32 floats; KU = number of 32-lane vector slots actually used by `cols`
// (3 for a 768-wide bf16 row), so the gamma / beta / bias-gradient accumulators are 3 x KU x N registers.
template <typename T, int KU, int NP = 0>
__global__ void __launch_bounds__(128) ln_bwd_packed(const T* __restrict__ dy, const T* __restrict__ xsum,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const T* __restrict__ dx_extra,
                                                     T* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, LnBwdExtra ex, int64_t rows, int cols) {
  extern __shared__ float sm[];  // [nwarp][3][cols] slabs for the final flush
  constexpr int N = Vec16<T>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float dg[KU * N], db[KU * N], dsum[KU * N];
#pragma unroll
  for (int k = 0; k < KU * N; ++k) dg[k] = db[k] = dsum[k] = 0.0f;
  pdl_wait();
  pdl_trigger();
  T* dxd = reinterpret_cast<T*>(ex.dx_drop);
  const bool want_sum = ex.dxsum != nullptr || ex.partials != nullptr;
  const bool drop = ex.p > 0.0f;
  if (drop) ex.seed = seed_eff(ex.seed, ex.seed_ctr);
  const uint32_t thr = (uint32_t)(ex.p * 4294967296.0);
  const float inv_keep = drop ? 1.0f / (1.0f - ex.p) : 1.0f;
  for (int64_t row = blockIdx.x * (int64_t)nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    const float mu = mean[row], rs = rstd[row];
    const T* dyr = dy + row * cols;
    const T* xr = xsum + row * cols;
    const T* er = dx_extra ? dx_extra + row * cols : nullptr;
    T* dxr = dx + row * cols;
    T* ddr = dxd ? dxd + row * cols : nullptr;
    Vec16<T> a[KU], b[KU];
    if (ex.dy_parts == nullptr) {
#pragma unroll
      for (int k = 0; k < KU; ++k) {
        const int c = (k * 32 + lane) * N;
        if (c < cols) {
          a[k].load(dyr + c);
          b[k].load(xr + c);
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < KU; ++k) {
        const int c = (k * 32 + lane) * N;
        if (c < cols) {
          b[k].load(xr + c);
          float acc[N];
#pragma unroll
          for (int j = 0; j < N; ++j) acc[j] = 0.0f;
          if (NP > 0) {      // compile-time slab count: all slab loads of the row in flight at once
            float4 t[NP > 0 ? NP : 1][N / 4];
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              const float* pr = ex.dy_parts + (int64_t)i * ex.part_stride + row * cols + c;
#pragma unroll
              for (int j = 0; j < N / 4; ++j) t[i][j] = __ldg(reinterpret_cast<const float4*>(pr) + j);
            }
#pragma unroll
            for (int i = 0; i < NP; ++i) {
#pragma unroll
              for (int j = 0; j < N / 4; ++j) {
                acc[4 * j] += t[i][j].x; acc[4 * j + 1] += t[i][j].y; acc[4 * j + 2] += t[i][j].z; acc[4 * j + 3] += t[i][j].w;
              }
            }
          } else {
            for (int i = 0; i < ex.nparts; ++i) {
              const float* pr = ex.dy_parts + (int64_t)i * ex.part_stride + row * cols + c;
#pragma unroll
              for (int j = 0; j < N; j += 4) {
                const float4 t = *reinterpret_cast<const float4*>(pr + j);
                acc[j] += t.x; acc[j + 1] += t.y; acc[j + 2] += t.z; acc[j + 3] += t.w;
              }
            }
          }
          if (dy) {
            Vec16<T> r;
            r.load(dyr + c);
#pragma unroll
            for (int j = 0; j < N; ++j) acc[j] += r.get(j);
          }
#pragma unroll
          for (int j = 0; j < N; ++j) a[k].set(j, acc[j]);
        }
      }
    }
    float s1 = 0.0f, s2 = 0.0f;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float g = a[k].get(j), xh = (b[k].get(j) - mu) * rs;
          dg[k * N + j] += g * xh;
          db[k * N + j] += g;
          const float gg = g * __ldg(gamma + c + j);
          s1 += gg;
          s2 += gg * xh;
        }
      }
    }
    s1 = warp_sum(s1) / (float)cols;
    s2 = warp_sum(s2) / (float)cols;
#pragma unroll
    for (int k = 0; k < KU; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < cols) {
        Vec16<T> e, o, od;
        if (er) e.load(er + c);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float gg = a[k].get(j) * __ldg(gamma + c + j), xh = (b[k].get(j) - mu) * rs;
          o.set(j, rs * (gg - s1 - xh * s2) + (er ? e.get(j) : 0.0f));
          if (ddr || want_sum) {
            float d = o.get(j);   // the stored (rounded) value is what the branch sees
            if (drop) d = hash32(ex.seed, (uint64_t)(row * cols + c + j)) >= thr ? d * inv_keep : 0.0f;
            od.set(j, d);
            dsum[k * N + j] += od.get(j);
          }
        }
        o.store(dxr + c);
        if (ddr) od.store(ddr + c);
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && !want_sum && ex.partials == nullptr) return;
  float* slab = sm + (size_t)warp * 3 * cols;
#pragma unroll
  for (int k = 0; k < KU; ++k) {
    const int c = (k * 32 + lane) * N;
    if (c < cols) {
#pragma unroll
      for (int j = 0; j < N; j += 4) {
        *reinterpret_cast<float4*>(slab + c + j) = make_float4(dg[k * N + j], dg[k * N + j + 1], dg[k * N + j + 2], dg[k * N + j + 3]);
        *reinterpret_cast<float4*>(slab + cols + c + j) = make_float4(db[k * N + j], db[k * N + j + 1], db[k * N + j + 2], db[k * N + j + 3]);
        *reinterpret_cast<float4*>(slab + 2 * cols + c + j) = make_float4(dsum[k * N + j], dsum[k * N + j + 1], dsum[k * N + j + 2], dsum[k * N + j + 3]);
      }
    }
  }
  __syncthreads();
  const int nv = cols / 4;
  for (int i = threadIdx.x; i < 3 * nv; i += blockDim.x) {
    const int which = i / nv, c4 = (i - which * nv) * 4;
    float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : ex.dxsum);
    if (dst == nullptr && ex.partials == nullptr) continue;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int w = 0; w < nwarp; ++w) {
      const float4 t = *reinterpret_cast<const float4*>(sm + (size_t)w * 3 * cols + which * cols + c4);
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    if (ex.partials) *reinterpret_cast<float4*>(ex.partials + ((size_t)blockIdx.x * 3 + which) * cols + c4) = acc;
    else atomicAdd(reinterpret_cast<float4*>(dst + c4), acc);
  }
}

// out[which][c] += sum_p partials[p][which][c]: the deferred half of ln_bwd_packed (partials mode), off the critical path.
// grid (ceil(3 * cols / 4 / 128), PSPLIT): a thread owns one float4 column group and a slice of the partial rows.
__global__ void __launch_bounds__(128) ln_partials_reduce_kernel(const float* __restrict__ partials, int nparts, int cols,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                 float* __restrict__ dxsum) {
  const int nv = cols / 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * nv) return;
  const int which = i / nv, c4 = (i - which * nv) * 4;
  float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : dxsum);
  if (dst == nullptr) return;
  const int per = (nparts + gridDim.y - 1) / gridDim.y;
  const int p0 = blockIdx.y * per, p1 = min(nparts, p0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int p = p0; p < p1; ++p) {
    const float4 t = *reinterpret_cast<const float4*>(partials + ((size_t)p * 3 + which) * cols + c4);
    acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
  }
  if (p1 > p0) atomicAdd(reinterpret_cast<float4*>(dst + c4), acc);
}

// backward.  Each warp walks rows (grid-stride); lanes own fixed columns, so dgamma/dbeta (and the column
// sums of the optional dropped copy) accumulate in registers and are flushed through shared memory with one
// global atomic per (block, column).  CACHED = cols <= 1024.
template <typename T, bool VEC, bool CACHED>
__global__ void __launch_bounds__(128) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ xsum,
                                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, const T* __restrict__ dx_extra,
                                                     T* __restrict__ dx, float* __restrict__ dgamma,
                                                     float* __restrict__ dbeta, LnBwdExtra ex, int64_t rows, int cols) {
  extern __shared__ float sm[];  // [3][cols] block partials
  constexpr int N = VEC ? Vec16<T>::N : 1;
  constexpr int ITER = LN_CACHE / N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float dg[CACHED ? LN_CACHE : 1], db[CACHED ? LN_CACHE : 1], dsum[CACHED ? LN_CACHE : 1];
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < LN_CACHE; ++k) dg[k] = db[k] = dsum[k] = 0.0f;
  }
  for (int c = threadIdx.x; c < 3 * cols; c += blockDim.x) sm[c] = 0.0f;
  __syncthreads();
  pdl_wait();
  pdl_trigger();
  T* dxd = reinterpret_cast<T*>(ex.dx_drop);
  const bool want_sum = ex.dxsum != nullptr;
  const bool drop = ex.p > 0.0f;
  if (drop) ex.seed = seed_eff(ex.seed, ex.seed_ctr);
  const uint32_t thr = (uint32_t)(ex.p * 4294967296.0);
  const float inv_keep = drop ? 1.0f / (1.0f - ex.p) : 1.0f;
  for (int64_t row = blockIdx.x * (int64_t)nwarp + warp; row < rows; row += (int64_t)gridDim.x * nwarp) {
    const float mu = mean[row], rs = rstd[row];
    const T* dyr = dy + row * cols;
    const T* xr = xsum + row * cols;
    const T* er = dx_extra ? dx_extra + row * cols : nullptr;
    T* dxr = dx + row * cols;
    T* ddr = dxd ? dxd + row * cols : nullptr;
    float s1 = 0.0f, s2 = 0.0f;
    if (CACHED) {
      float g_[LN_CACHE], xh_[LN_CACHE];
#pragma unroll
      for (int k = 0; k < ITER; ++k) {
        const int c = (k * 32 + lane) * N;
        if (c < cols) {
          if (VEC) {
            Vec16<T> a, b;
            a.load(dyr + c);
            b.load(xr + c);
#pragma unroll
            for (int j = 0; j < N; ++j) {
              float g = a.get(j), xh = (b.get(j) - mu) * rs;
              dg[k * N + j] += g * xh;
              db[k * N + j] += g;
              g *= __ldg(gamma + c + j);
              g_[k * N + j] = g;
              xh_[k * N + j] = xh;
              s1 += g;
              s2 += g * xh;
            }
          } else {
            float g = to_f(dyr[c]), xh = (to_f(xr[c]) - mu) * rs;
            dg[k] += g * xh;
            db[k] += g;
            g *= __ldg(gamma + c);
            g_[k] = g;
            xh_[k] = xh;
            s1 += g;
            s2 += g * xh;
          }
        }
      }
      s1 = warp_sum(s1) / (float)cols;
      s2 = warp_sum(s2) / (float)cols;
#pragma unroll
      for (int k = 0; k < ITER; ++k) {
        const int c = (k * 32 + lane) * N;
        if (c < cols) {
          if (VEC) {
            Vec16<T> e, o, od;
            if (er) e.load(er + c);
#pragma unroll
            for (int j = 0; j < N; ++j) {
              o.set(j, rs * (g_[k * N + j] - s1 - xh_[k * N + j] * s2) + (er ? e.get(j) : 0.0f));
              if (ddr || want_sum) {
                float d = o.get(j);   // the stored (rounded) value is what the branch sees
                if (drop) d = hash32(ex.seed, (uint64_t)(row * cols + c + j)) >= thr ? d * inv_keep : 0.0f;
                od.set(j, d);
                dsum[k * N + j] += od.get(j);
              }
            }
            o.store(dxr + c);
            if (ddr) od.store(ddr + c);
          } else {
            T o = from_f<T>(rs * (g_[k] - s1 - xh_[k] * s2) + (er ? to_f(er[c]) : 0.0f));
            dxr[c] = o;
            if (ddr || want_sum) {
              float d = to_f(o);
              if (drop) d = hash32(ex.seed, (uint64_t)(row * cols + c)) >= thr ? d * inv_keep : 0.0f;
              T q = from_f<T>(d);
              if (ddr) ddr[c] = q;
              dsum[k] += to_f(q);
            }
          }
        }
      }
    } else {
      for (int c = lane; c < cols; c += 32) {
        float g = to_f(dyr[c]), xh = (to_f(xr[c]) - mu) * rs;
        atomicAdd(&sm[c], g * xh);
        atomicAdd(&sm[cols + c], g);
        g *= __ldg(gamma + c);
        s1 += g;
        s2 += g * xh;
      }
      s1 = warp_sum(s1) / (float)cols;
      s2 = warp_sum(s2) / (float)cols;
      for (int c = lane; c < cols; c += 32) {
        float g = to_f(dyr[c]) * __ldg(gamma + c), xh = (to_f(xr[c]) - mu) * rs;
        T o = from_f<T>(rs * (g - s1 - xh * s2) + (er ? to_f(er[c]) : 0.0f));
        dxr[c] = o;
        if (ddr || want_sum) {
          float d = to_f(o);
          if (drop) d = hash32(ex.seed, (uint64_t)(row * cols + c)) >= thr ? d * inv_keep : 0.0f;
          T q = from_f<T>(d);
          if (ddr) ddr[c] = q;
          atomicAdd(&sm[2 * cols + c], to_f(q));
        }
      }
    }
  }
  if (dgamma == nullptr && dbeta == nullptr && !want_sum) return;
  if (CACHED && VEC) {
    // each warp parks its register partials in a private shared-memory slab (plain 128-bit stores, no atomics),
    // then the CTA sums the slabs and issues ONE 128-bit vector atomic per 4 columns and output
    __syncthreads();                      // sm[] was only used as zero-filled scratch so far
    float* slab = sm + (size_t)warp * 3 * cols;
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          *reinterpret_cast<float4*>(slab + c + j) = make_float4(dg[k * N + j], dg[k * N + j + 1], dg[k * N + j + 2], dg[k * N + j + 3]);
          *reinterpret_cast<float4*>(slab + cols + c + j) = make_float4(db[k * N + j], db[k * N + j + 1], db[k * N + j + 2], db[k * N + j + 3]);
          if (want_sum)
            *reinterpret_cast<float4*>(slab + 2 * cols + c + j) = make_float4(dsum[k * N + j], dsum[k * N + j + 1], dsum[k * N + j + 2], dsum[k * N + j + 3]);
        }
      }
    }
    __syncthreads();
    const int nv = cols / 4;
    for (int i = threadIdx.x; i < 3 * nv; i += blockDim.x) {
      const int which = i / nv, c4 = (i - which * nv) * 4;
      float* dst = which == 0 ? dgamma : (which == 1 ? dbeta : ex.dxsum);
      if (dst == nullptr) continue;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int w = 0; w < nwarp; ++w) {
        const float4 t = *reinterpret_cast<const float4*>(sm + (size_t)w * 3 * cols + which * cols + c4);
        acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
      }
      atomicAdd(reinterpret_cast<float4*>(dst + c4), acc);
    }
    return;
  }
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < ITER; ++k) {
      const int c = (k * 32 + lane) * N;
      if (c < cols) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          atomicAdd(&sm[c + j], dg[k * N + j]);
          atomicAdd(&sm[cols + c + j], db[k * N + j]);
          if (want_sum) atomicAdd(&sm[2 * cols + c + j], dsum[k * N + j]);
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + c, sm[c]);
    if (dbeta) atomicAdd(dbeta + c, sm[cols + c]);
    if (want_sum) atomicAdd(ex.dxsum + c, sm[2 * cols + c]);
  }
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_abi_version(void) { return MMVQA_ABI_VERSION; }
const char* mmvqa_last_error(void) { return g_err; }
int64_t mmvqa_launch_count(void) { return g_launches.load(); }
int mmvqa_set_dropout_counter(const uint64_t* counter) {
  g_seed_ctr = reinterpret_cast<const unsigned long long*>(counter);
  return MMVQA_OK;
}

int mmvqa_device_sm(void) {
  int dev = 0, major = 0, minor = 0;
  MMVQA_CUDA(cudaGetDevice(&dev));
  MMVQA_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  MMVQA_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}

int mmvqa_bias_act_fwd(const void* x, const float* bias, void* y, int64_t rows, int cols, int act, int dtype,
                       mmvqa_stream_t stream) {
  MMVQA_REQUIRE(x && y, "bias_act_fwd: null pointer");
  MMVQA_REQUIRE(act >= MMVQA_ACT_NONE && act <= MMVQA_ACT_RELU, "bias_act_fwd: bad act %d", act);
  if (dtype == MMVQA_F32) return bias_act_launch<float, false>(x, bias, nullptr, y, rows, cols, act, as_stream(stream));
  if (dtype == MMVQA_BF16)
    return bias_act_launch<__nv_bfloat16, false>(x, bias, nullptr, y, rows, cols, act, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "bias_act_fwd: bad dtype %d", dtype);
}

int mmvqa_bias_act_bwd(const void* x, const float* bias, const void* dy, void* dx, int64_t rows, int cols, int act,
                       int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(x && dy && dx, "bias_act_bwd: null pointer");
  MMVQA_REQUIRE(act >= MMVQA_ACT_NONE && act <= MMVQA_ACT_RELU, "bias_act_bwd: bad act %d", act);
  if (dtype == MMVQA_F32) return bias_act_launch<float, true>(x, bias, dy, dx, rows, cols, act, as_stream(stream));
  if (dtype == MMVQA_BF16) return bias_act_launch<__nv_bfloat16, true>(x, bias, dy, dx, rows, cols, act, as_stream(stream));
  return set_err(MMVQA_ERR_ARG, "bias_act_bwd: bad dtype %d", dtype);
}

int mmvqa_l2_prefetch(const mmvqa_prefetch_list* list, int ctas, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(list && list->n >= 0 && list->n <= MMVQA_PREFETCH_MAX, "l2_prefetch: bad list");
  if (list->n == 0) return MMVQA_OK;
  for (int i = 0; i < list->n; ++i) MMVQA_REQUIRE(list->ptr[i] && list->bytes[i] >= 0, "l2_prefetch: bad range %d", i);
  if (ctas <= 0) ctas = 16;
  l2_prefetch_kernel<<<ctas, 256, 0, as_stream(stream)>>>(*list);
  MMVQA_LAUNCHED("l2_prefetch");
  return MMVQA_OK;
}

int mmvqa_colsum(const void* x, int64_t ldx, float* out, int64_t rows, int cols, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(out && (x || rows == 0), "colsum: null pointer");
  MMVQA_REQUIRE(cols > 0 && ldx >= cols, "colsum: bad cols/ld");
  cudaStream_t st = as_stream(stream);
  if (rows <= 0) {
    MMVQA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, st));
    return MMVQA_OK;
  }
  int gx = (cols + 31) / 32;
  int64_t want = ((int64_t)num_sms() * 4 + gx - 1) / gx;  // row splits to fill the chip
  int64_t rpb = (rows + want - 1) / want;
  if (rpb < 512) rpb = 512;
  int gy = (int)((rows + rpb - 1) / rpb);
  if (gy > 1) MMVQA_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)cols, st));
  dim3 grid(gx, gy);
  if (dtype == MMVQA_F32)
    colsum_kernel<float><<<grid, 256, 0, st>>>((const float*)x, ldx, out, rows, cols, rpb);
  else if (dtype == MMVQA_BF16)
    colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, ldx, out, rows, cols, rpb);
  else
    return set_err(MMVQA_ERR_ARG, "colsum: bad dtype %d", dtype);
  MMVQA_LAUNCHED("colsum");
  return MMVQA_OK;
}

int mmvqa_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(n >= 0 && (n == 0 || (src && dst)), "cast: null pointer");
  if (n == 0) return MMVQA_OK;
  cudaStream_t st = as_stream(stream);
  int64_t maxg = (int64_t)num_sms() * 16;
  if (src_dtype == MMVQA_F32 && dst_dtype == MMVQA_BF16 && n % 4 == 0 && aligned16(src) &&
      (reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
    int64_t n4 = n / 4;
    int grid = (int)((n4 + 255) / 256 < maxg ? (n4 + 255) / 256 : maxg);
    cast_f32_bf16_vec<<<grid, 256, 0, st>>>((const float4*)src, (uint2*)dst, n4);
    MMVQA_LAUNCHED("cast_f32_bf16");
    return MMVQA_OK;
  }
  int64_t work = (n + 3) / 4;
  int grid = (int)((work + 255) / 256 < maxg ? (work + 255) / 256 : maxg);
  if (src_dtype == MMVQA_F32 && dst_dtype == MMVQA_BF16)
    cast_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == MMVQA_BF16 && dst_dtype == MMVQA_F32)
    cast_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == MMVQA_F32 && dst_dtype == MMVQA_F32)
    cast_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, (float*)dst, n);
  else if (src_dtype == MMVQA_BF16 && dst_dtype == MMVQA_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, n);
  else
    return set_err(MMVQA_ERR_ARG, "cast: bad dtypes %d -> %d", src_dtype, dst_dtype);
  MMVQA_LAUNCHED("cast");
  return MMVQA_OK;
}

int mmvqa_cast_pad(const void* src, int src_dtype, int64_t ld_src, void* dst, int dst_dtype, int64_t ld_dst,
                   int64_t rows, int cols, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(src && dst, "cast_pad: null pointer");
  MMVQA_REQUIRE(ld_src >= cols && ld_dst >= cols && rows >= 0, "cast_pad: bad shape");
  if (rows == 0) return MMVQA_OK;
  cudaStream_t st = as_stream(stream);
  int64_t total = rows * ld_dst, maxg = (int64_t)num_sms() * 16;
  int grid = (int)((total + 255) / 256 < maxg ? (total + 255) / 256 : maxg);
  if (src_dtype == MMVQA_F32 && dst_dtype == MMVQA_BF16)
    cast_pad_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols);
  else if (src_dtype == MMVQA_BF16 && dst_dtype == MMVQA_F32)
    cast_pad_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, ld_src, (float*)dst, ld_dst, rows, cols);
  else if (src_dtype == MMVQA_F32 && dst_dtype == MMVQA_F32)
    cast_pad_kernel<float, float><<<grid, 256, 0, st>>>((const float*)src, ld_src, (float*)dst, ld_dst, rows, cols);
  else if (src_dtype == MMVQA_BF16 && dst_dtype == MMVQA_BF16)
    cast_pad_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)src, ld_src, (__nv_bfloat16*)dst, ld_dst, rows, cols);
  else
    return set_err(MMVQA_ERR_ARG, "cast_pad: bad dtypes %d -> %d", src_dtype, dst_dtype);
  MMVQA_LAUNCHED("cast_pad");
  return MMVQA_OK;
}

int mmvqa_cast_pad_multi(const mmvqa_cast_list* list, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(list && list->n >= 0 && list->n <= MMVQA_CAST_MULTI_MAX, "cast_pad_multi: bad list");
  if (list->n == 0) return MMVQA_OK;
  CastMultiArgs a;
  a.l = *list;
  int64_t run = 0;
  for (int i = 0; i < list->n; ++i) {
    MMVQA_REQUIRE(list->src[i] && list->dst[i], "cast_pad_multi: null pointer in problem %d", i);
    MMVQA_REQUIRE(list->rows[i] >= 0 && list->cols[i] > 0 && list->ld_src[i] >= list->cols[i] && list->ld_dst[i] >= list->cols[i],
                  "cast_pad_multi: bad shape in problem %d", i);
    MMVQA_REQUIRE(list->ld_dst[i] % 4 == 0 && (reinterpret_cast<uintptr_t>(list->dst[i]) & 7) == 0,
                  "cast_pad_multi: problem %d needs ld_dst %% 4 == 0 and an 8-byte aligned destination", i);
    run += list->rows[i] * (list->ld_dst[i] / 4);
    a.unit_end[i] = run;
    a.vec_src[i] = (list->ld_src[i] % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(list->src[i]) & (list->src_bf16[i] ? 7 : 15)) == 0) ? 1 : 0;
  }
  for (int i = list->n; i < MMVQA_CAST_MULTI_MAX; ++i) { a.unit_end[i] = run; a.vec_src[i] = 0; }
  if (run == 0) return MMVQA_OK;
  const int64_t maxg = (int64_t)num_sms() * 8, want = (run + 255) / 256;
  cast_pad_multi_kernel<<<(int)(want < maxg ? want : maxg), 256, 0, as_stream(stream)>>>(a);
  MMVQA_LAUNCHED("cast_pad_multi");
  return MMVQA_OK;
}

int mmvqa_scale_by_device_scalar(void* x, int dtype, const float* scalar, float host_factor, int64_t n,
                                 mmvqa_stream_t stream) {
  MMVQA_REQUIRE(n >= 0 && (n == 0 || x), "scale: null pointer");
  if (n == 0) return MMVQA_OK;
  int64_t maxg = (int64_t)num_sms() * 16;
  int grid = (int)((n + 255) / 256 < maxg ? (n + 255) / 256 : maxg);
  if (dtype == MMVQA_F32)
    scale_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((float*)x, scalar, host_factor, n);
  else if (dtype == MMVQA_BF16)
    scale_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((__nv_bfloat16*)x, scalar, host_factor, n);
  else
    return set_err(MMVQA_ERR_ARG, "scale: bad dtype %d", dtype);
  MMVQA_LAUNCHED("scale");
  return MMVQA_OK;
}

int mmvqa_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(n >= 0 && (n == 0 || (x && y)), "dropout: null pointer");
  MMVQA_REQUIRE(p >= 0.0f && p < 1.0f, "dropout: p must be in [0,1)");
  if (n == 0) return MMVQA_OK;
  int64_t maxg = (int64_t)num_sms() * 16;
  int grid = (int)((n + 255) / 256 < maxg ? (n + 255) / 256 : maxg);
  if (dtype == MMVQA_F32)
    dropout_kernel<float><<<grid, 256, 0, as_stream(stream)>>>((const float*)x, (float*)y, n, p, seed, g_seed_ctr);
  else if (dtype == MMVQA_BF16)
    dropout_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, p, seed, g_seed_ctr);
  else
    return set_err(MMVQA_ERR_ARG, "dropout: bad dtype %d", dtype);
  MMVQA_LAUNCHED("dropout");
  return MMVQA_OK;
}

int mmvqa_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                            void* sum_out, float* mean, float* rstd, int64_t rows, int cols, float eps, int dtype,
                            mmvqa_stream_t stream) {
  MMVQA_REQUIRE(x && gamma && beta && y, "add_layernorm_fwd: null pointer");
  MMVQA_REQUIRE(cols > 0 && rows >= 0, "add_layernorm_fwd: bad shape");
  MMVQA_REQUIRE(dtype == MMVQA_F32 || dtype == MMVQA_BF16, "add_layernorm_fwd: bad dtype %d", dtype);
  if (rows == 0) return MMVQA_OK;
  cudaStream_t st = as_stream(stream);
  int grid = (int)((rows + 3) / 4);
  const int vn = dtype == MMVQA_F32 ? 4 : 8;
  const bool cached = cols <= 32 * LN_CACHE;
  const bool vec = cached && cols % vn == 0 && aligned16(x) && aligned16(y) && (!res || aligned16(res)) &&
                   (!sum_out || aligned16(sum_out));
#define LN_FWD(T, K) launch_pdl(K, dim3(grid), dim3(128), 0, st, (const T*)x, (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, rstd, rows, cols, eps)
#define LN_FWD_PACKED(T) launch_pdl(add_ln_fwd_packed<T>, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, st, (const T*)x, (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, rstd, rows, cols, eps)
  const bool gb_al = aligned16(gamma) && aligned16(beta);
  if (vec && gb_al) {
    if (dtype == MMVQA_F32) LN_FWD_PACKED(float);
    else LN_FWD_PACKED(__nv_bfloat16);
  } else if (dtype == MMVQA_F32) {
    if (vec) LN_FWD(float, (add_ln_fwd_kernel<float, true>));
    else if (cached) LN_FWD(float, (add_ln_fwd_kernel<float, false>));
    else LN_FWD(float, add_ln_fwd_generic<float>);
  } else {
    using B = __nv_bfloat16;
    if (vec) LN_FWD(B, (add_ln_fwd_kernel<B, true>));
    else if (cached) LN_FWD(B, (add_ln_fwd_kernel<B, false>));
    else LN_FWD(B, add_ln_fwd_generic<B>);
  }
#undef LN_FWD
#undef LN_FWD_PACKED
  MMVQA_LAUNCHED("add_layernorm_fwd");
  return MMVQA_OK;
}

// CTAs of the packed backward kernel for this problem (= rows of the partials workspace), 0 if it does not apply
static int ln_bwd_packed_grid(int64_t rows, int cols, int dtype) {
  const int vn = dtype == MMVQA_F32 ? 4 : 8;
  if (rows <= 0 || cols <= 0 || cols > 32 * LN_CACHE || cols % vn != 0) return 0;
  const int nw = 4;
  if (sizeof(float) * 3 * (size_t)cols * nw > 48 * 1024) return 0;
  int64_t want2 = (rows + nw - 1) / nw, cap2 = (int64_t)num_sms() * 12;
  return (int)(want2 < cap2 ? want2 : cap2);
}

static int layernorm_bwd_impl(const void* dy, const float* dy_parts, int nparts, int64_t part_stride, const void* xsum,
                              const float* gamma, const float* mean, const float* rstd, const void* dx_extra, void* dx,
                              float* dgamma, float* dbeta, void* dx_drop, float* dxsum, float dropout_p,
                              uint64_t dropout_seed, int64_t rows, int cols, int dtype, mmvqa_stream_t stream,
                              float* partials = nullptr, int partial_rows = 0) {
  MMVQA_REQUIRE((dy || dy_parts) && xsum && gamma && mean && rstd && dx, "layernorm_bwd: null pointer");
  MMVQA_REQUIRE(cols > 0 && rows >= 0, "layernorm_bwd: bad shape");
  MMVQA_REQUIRE(dtype == MMVQA_F32 || dtype == MMVQA_BF16, "layernorm_bwd: bad dtype %d", dtype);
  MMVQA_REQUIRE(dropout_p >= 0.0f && dropout_p < 1.0f, "layernorm_bwd: dropout_p must be in [0,1)");
  if (rows == 0) return MMVQA_OK;
  cudaStream_t st = as_stream(stream);
  int64_t want = (rows + 3) / 4, cap = (int64_t)num_sms() * 4;
  int grid = (int)(want < cap ? want : cap);
  const int vn = dtype == MMVQA_F32 ? 4 : 8;
  const bool cached = cols <= 32 * LN_CACHE;
  const bool vec = cached && cols % vn == 0 && (!dy || aligned16(dy)) && aligned16(xsum) && aligned16(dx) &&
                   (!dx_extra || aligned16(dx_extra)) && (!dx_drop || aligned16(dx_drop)) &&
                   (!dgamma || aligned16(dgamma)) && (!dbeta || aligned16(dbeta)) && (!dxsum || aligned16(dxsum)) &&
                   (!dy_parts || (aligned16(dy_parts) && part_stride % 4 == 0));
  size_t smem = sizeof(float) * 3 * (size_t)cols * (vec ? 4 : 1);     // vec path: one slab per warp
  MMVQA_REQUIRE(smem <= 48 * 1024, "layernorm_bwd: cols %d too large", cols);
  LnBwdExtra ex;
  ex.dx_drop = dx_drop; ex.dxsum = dxsum; ex.p = dx_drop ? dropout_p : 0.0f; ex.seed = dropout_seed; ex.seed_ctr = g_seed_ctr;
  ex.dy_parts = dy_parts; ex.nparts = nparts; ex.part_stride = part_stride;
  ex.partials = partials;
  if (partials) {
    MMVQA_REQUIRE(vec && aligned16(gamma) && aligned16(partials) && partial_rows == ln_bwd_packed_grid(rows, cols, dtype) &&
                  partial_rows > 0, "layernorm_bwd: the partials mode needs the packed kernel (aligned, cols %% %d == 0) and "
                  "a workspace of mmvqa_layernorm_bwd_partial_rows() rows", vn);
  }
  // packed kernel: 4-warp CTAs (one row per warp) when the problem is small, 8-warp CTAs with a row loop otherwise
  if (vec && aligned16(gamma)) {
    const int vnn = dtype == MMVQA_F32 ? 4 : 8;
    const int ku = (cols + 32 * vnn - 1) / (32 * vnn);
    const int nthr = 128;
    const int nw = nthr / 32;
    int64_t want2 = (rows + nw - 1) / nw, cap2 = (int64_t)num_sms() * 12;
    const int grid2 = (int)(want2 < cap2 ? want2 : cap2);
    const size_t smem2 = sizeof(float) * 3 * (size_t)cols * nw;
    bool launched = true;
    if (smem2 > 48 * 1024) launched = false;
#define LN_BWDP(T, KU) MMVQA_CUDA(launch_pdl(ln_bwd_packed<T, KU>, dim3(grid2), dim3(nthr), smem2, st, (const T*)dy, (const T*)xsum, gamma, mean, rstd, (const T*)dx_extra, (T*)dx, dgamma, dbeta, ex, rows, cols))
    if (launched) {
      if (dtype == MMVQA_BF16) {
        using B = __nv_bfloat16;
        // (three slabs = the K = 3072 dgrad at M = 448 on 148 SMs: compile-time slab count for the 768-wide row)
        if (ku == 3 && dy_parts != nullptr && nparts == 3)
          MMVQA_CUDA(launch_pdl(ln_bwd_packed<B, 3, 3>, dim3(grid2), dim3(nthr), smem2, st, (const B*)dy, (const B*)xsum, gamma, mean, rstd, (const B*)dx_extra, (B*)dx, dgamma, dbeta, ex, rows, cols));
        else if (ku <= 1) LN_BWDP(B, 1); else if (ku == 2) LN_BWDP(B, 2); else if (ku == 3) LN_BWDP(B, 3); else LN_BWDP(B, 4);
      } else {
        if (ku <= 2) LN_BWDP(float, 2); else if (ku <= 4) LN_BWDP(float, 4); else if (ku <= 6) LN_BWDP(float, 6); else LN_BWDP(float, 8);
      }
      MMVQA_LAUNCHED("layernorm_bwd");
      return MMVQA_OK;
    }
#undef LN_BWDP
  }
  MMVQA_REQUIRE(dy_parts == nullptr, "layernorm_bwd_parts: needs cols %% %d == 0, cols <= 1024 and 16-byte aligned buffers", vn);
#define LN_BWD(T, V, C) launch_pdl(ln_bwd_kernel<T, V, C>, dim3(grid), dim3(128), smem, st, (const T*)dy, (const T*)xsum, gamma, mean, rstd, (const T*)dx_extra, (T*)dx, dgamma, dbeta, ex, rows, cols)
  if (dtype == MMVQA_F32) {
    if (vec) LN_BWD(float, true, true);
    else if (cached) LN_BWD(float, false, true);
    else LN_BWD(float, false, false);
  } else {
    using B = __nv_bfloat16;
    if (vec) LN_BWD(B, true, true);
    else if (cached) LN_BWD(B, false, true);
    else LN_BWD(B, false, false);
  }
#undef LN_BWD
  MMVQA_LAUNCHED("layernorm_bwd");
  return MMVQA_OK;
}

int mmvqa_layernorm_bwd(const void* dy, const void* xsum, const float* gamma, const float* mean, const float* rstd,
                        const void* dx_extra, void* dx, float* dgamma, float* dbeta, void* dx_drop, float* dxsum,
                        float dropout_p, uint64_t dropout_seed, int64_t rows, int cols, int dtype,
                        mmvqa_stream_t stream) {
  MMVQA_REQUIRE(dy != nullptr, "layernorm_bwd: null pointer");
  return layernorm_bwd_impl(dy, nullptr, 0, 0, xsum, gamma, mean, rstd, dx_extra, dx, dgamma, dbeta, dx_drop, dxsum,
                            dropout_p, dropout_seed, rows, cols, dtype, stream);
}

int mmvqa_layernorm_bwd_parts(const float* dy_parts, int nparts, int64_t part_stride, const void* dy_res,
                              const void* xsum, const float* gamma, const float* mean, const float* rstd, void* dx,
                              float* dgamma, float* dbeta, void* dx_drop, float* dxsum, float dropout_p,
                              uint64_t dropout_seed, int64_t rows, int cols, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(dy_parts != nullptr && nparts >= 1 && part_stride >= rows * cols, "layernorm_bwd_parts: bad partials");
  return layernorm_bwd_impl(dy_res, dy_parts, nparts, part_stride, xsum, gamma, mean, rstd, nullptr, dx, dgamma, dbeta,
                            dx_drop, dxsum, dropout_p, dropout_seed, rows, cols, dtype, stream);
}

int mmvqa_layernorm_bwd_partial_rows(int64_t rows, int cols, int dtype) { return ln_bwd_packed_grid(rows, cols, dtype); }

int mmvqa_layernorm_bwd_deferred(const void* dy, const float* dy_parts, int nparts, int64_t part_stride, const void* xsum,
                                 const float* gamma, const float* mean, const float* rstd, void* dx, void* dx_drop,
                                 float dropout_p, uint64_t dropout_seed, int64_t rows, int cols, int dtype, float* partials,
                                 int partial_rows, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(partials != nullptr, "layernorm_bwd_deferred: null workspace");
  MMVQA_REQUIRE(dy_parts == nullptr || (nparts >= 1 && part_stride >= rows * cols), "layernorm_bwd_deferred: bad partial tiles");
  return layernorm_bwd_impl(dy, dy_parts, nparts, part_stride, xsum, gamma, mean, rstd, nullptr, dx, nullptr, nullptr, dx_drop,
                            nullptr, dropout_p, dropout_seed, rows, cols, dtype, stream, partials, partial_rows);
}

int mmvqa_ln_partials_reduce(const float* partials, int partial_rows, int cols, float* dgamma, float* dbeta, float* dxsum,
                             mmvqa_stream_t stream) {
  MMVQA_REQUIRE(partials && partial_rows > 0 && cols > 0 && cols % 4 == 0, "ln_partials_reduce: bad args");
  MMVQA_REQUIRE(aligned16(partials) && (!dgamma || aligned16(dgamma)) && (!dbeta || aligned16(dbeta)) && (!dxsum || aligned16(dxsum)),
                "ln_partials_reduce: 16-byte aligned buffers needed");
  const int nv = 3 * cols / 4;
  const int psplit = partial_rows >= 32 ? 4 : 1;
  ln_partials_reduce_kernel<<<dim3((nv + 127) / 128, psplit), 128, 0, as_stream(stream)>>>(partials, partial_rows, cols, dgamma,
                                                                                           dbeta, dxsum);
  MMVQA_LAUNCHED("ln_partials_reduce");
  return MMVQA_OK;
}

int mmvqa_add_layernorm_fwd_parts(const float* parts, int nparts, int64_t part_stride, const void* res,
                                  const float* gamma, const float* beta, void* y, void* sum_out, float* mean,
                                  float* rstd, int64_t rows, int cols, float eps, float dropout_p,
                                  uint64_t dropout_seed, int dtype, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(parts && gamma && beta && y, "add_layernorm_fwd_parts: null pointer");
  MMVQA_REQUIRE(nparts >= 1 && part_stride >= rows * cols && part_stride % 4 == 0, "add_layernorm_fwd_parts: bad partials");
  MMVQA_REQUIRE(cols > 0 && rows >= 0, "add_layernorm_fwd_parts: bad shape");
  MMVQA_REQUIRE(dtype == MMVQA_F32 || dtype == MMVQA_BF16, "add_layernorm_fwd_parts: bad dtype %d", dtype);
  MMVQA_REQUIRE(dropout_p >= 0.0f && dropout_p < 1.0f, "add_layernorm_fwd_parts: dropout_p must be in [0,1)");
  if (rows == 0) return MMVQA_OK;
  const int vn = dtype == MMVQA_F32 ? 4 : 8;
  MMVQA_REQUIRE(cols <= 32 * LN_CACHE && cols % vn == 0 && aligned16(parts) && aligned16(y) && (!res || aligned16(res)) &&
                    (!sum_out || aligned16(sum_out)) && aligned16(gamma) && aligned16(beta),
                "add_layernorm_fwd_parts: needs cols %% %d == 0, cols <= %d and 16-byte aligned buffers", vn, 32 * LN_CACHE);
  cudaStream_t st = as_stream(stream);
  const dim3 grid((unsigned)((rows + 7) / 8));
#define LN_PARTS(T, NP) MMVQA_CUDA(launch_pdl(add_ln_fwd_parts<T, NP>, grid, dim3(256), 0, st, parts, nparts, part_stride, (const T*)res, gamma, beta, (T*)y, (T*)sum_out, mean, rstd, rows, cols, eps, dropout_p, (unsigned long long)dropout_seed, g_seed_ctr))
  if (dtype == MMVQA_F32) {
    if (nparts == 2) LN_PARTS(float, 2); else if (nparts == 3) LN_PARTS(float, 3); else if (nparts == 4) LN_PARTS(float, 4); else LN_PARTS(float, 0);
  } else {
    using B = __nv_bfloat16;
    if (nparts == 2) LN_PARTS(B, 2); else if (nparts == 3) LN_PARTS(B, 3); else if (nparts == 4) LN_PARTS(B, 4); else LN_PARTS(B, 0);
  }
#undef LN_PARTS
  MMVQA_LAUNCHED("add_layernorm_fwd_parts");
  return MMVQA_OK;
}

}  // extern "C"
