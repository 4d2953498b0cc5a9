// common.cuh -- shared helpers for libmmvqa_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/mmvqa.h"

namespace mmvqa {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;
// device-resident dropout step counter (mmvqa_set_dropout_counter): every kernel that draws a dropout mask mixes
// *g_seed_ctr into its seed, so a CUDA graph that increments the counter gets fresh masks on every replay
extern const unsigned long long* g_seed_ctr;

int set_err(int code, const char* fmt, ...);

#define MMVQA_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::mmvqa::set_err(MMVQA_ERR_ARG, __VA_ARGS__); \
  } while (0)

// call after every kernel launch: counts it and converts launch errors
#define MMVQA_LAUNCHED(name)                                                            \
  do {                                                                                  \
    ::mmvqa::g_launches.fetch_add(1, std::memory_order_relaxed);                        \
    cudaError_t e__ = cudaPeekAtLastError();                                            \
    if (e__ != cudaSuccess) {                                                           \
      cudaGetLastError();                                                               \
      return ::mmvqa::set_err(MMVQA_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
    }                                                                                   \
  } while (0)

#define MMVQA_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) return ::mmvqa::set_err(MMVQA_ERR_CUDA, #call ": %s", cudaGetErrorString(e__)); \
  } while (0)

static inline cudaStream_t as_stream(mmvqa_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int num_sms();
bool pdl_enabled();

// Programmatic dependent launch: the kernel may be scheduled while the previous kernel on the stream is still
// draining; it must call pdl_wait() before touching any global memory a predecessor may have written (or that a
// predecessor may still be reading and this kernel overwrites).  Everything before pdl_wait() -- shared-memory
// carve-up, mbarrier init, TMEM allocation, tensor-map prefetch -- overlaps the predecessor's tail.
// Also valid inside stream capture (becomes a programmatic edge of the CUDA graph).
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
// same launch without the programmatic edge: the kernel starts only when its predecessor has completed
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_plain(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                       Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cfg.numAttrs = 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif

// ---------------------------------------------------------------------------------
// element access templated on storage type
// ---------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 128-bit vector of T: 4 floats or 8 bf16
template <typename T> struct Vec16;
template <> struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ void load(const float* p) { raw = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void store(float* p) const { *reinterpret_cast<float4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = raw; }
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (&raw.x)[i >> 1];
    return __uint_as_float((i & 1) ? (w & 0xffff0000u) : (w << 16));
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
    uint32_t& w = (&raw.x)[i >> 1];
    w = (i & 1) ? ((w & 0x0000ffffu) | (b << 16)) : ((w & 0xffff0000u) | b);
  }
};

// ---------------------------------------------------------------------------------
// activations.  act codes: MMVQA_ACT_*
//   SERF  models/serf.py:23-24   x * erf(log1p(exp(min(x, 50))))
//   GELU  models/transformer.py:7-8   exact erf form
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float serf_f(float x) {
  float sp = log1pf(expf(fminf(x, 50.0f)));
  return x * erff(sp);
}
// d/dx serf = erf(sp) + x * 2/sqrt(pi) * exp(-sp^2) * sigmoid(x)   (x <= 50);  erf(sp(50)) beyond the clamp
__device__ __forceinline__ float dserf_f(float x) {
  float xc = fminf(x, 50.0f);
  float sp = log1pf(expf(xc));
  float e = erff(sp);
  if (x > 50.0f) return e;
  float sig = 1.0f / (1.0f + expf(-x));
  return e + x * 1.1283791670955126f * expf(-sp * sp) * sig;
}
__device__ __forceinline__ float gelu_f(float x) { return x * 0.5f * (1.0f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float dgelu_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}
template <int ACT> __device__ __forceinline__ float act_f(float x) {
  if (ACT == MMVQA_ACT_SERF) return serf_f(x);
  if (ACT == MMVQA_ACT_GELU) return gelu_f(x);
  if (ACT == MMVQA_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}
template <int ACT> __device__ __forceinline__ float dact_f(float x) {
  if (ACT == MMVQA_ACT_SERF) return dserf_f(x);
  if (ACT == MMVQA_ACT_GELU) return dgelu_f(x);
  if (ACT == MMVQA_ACT_RELU) return x > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}
__device__ __forceinline__ float act_rt(int act, float x) {
  switch (act) {
    case MMVQA_ACT_SERF: return serf_f(x);
    case MMVQA_ACT_GELU: return gelu_f(x);
    case MMVQA_ACT_RELU: return fmaxf(x, 0.0f);
    default: return x;
  }
}
__device__ __forceinline__ float dact_rt(int act, float x) {
  switch (act) {
    case MMVQA_ACT_SERF: return dserf_f(x);
    case MMVQA_ACT_GELU: return dgelu_f(x);
    case MMVQA_ACT_RELU: return x > 0.0f ? 1.0f : 0.0f;
    default: return 1.0f;
  }
}

// ---------------------------------------------------------------------------------
// fast activations for the bf16 tensor-core epilogues (results are rounded to bf16 or pooled
// anyway): MUFU ex2/lg2/rcp based softplus and the Abramowitz-Stegun 7.1.26 erf
// (|abs err| < 1.5e-7 on erf, ~1e-6 relative on the activation).  The fp32 validation path
// keeps the accurate libdevice versions above.
// ---------------------------------------------------------------------------------
struct ErfPos {  // erf(z) for z >= 0 and E = exp(-z^2)
  float erf, E;
};
__device__ __forceinline__ ErfPos erf_pos_fast(float z) {
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  ErfPos r;
  r.E = __expf(-z * z);
  r.erf = fmaf(-poly, r.E, 1.0f);
  return r;
}
// erf(z), z >= 0, without an exponential: Abramowitz-Stegun 7.1.28, 1 - (1 + a1 z + ... + a6 z^6)^-16
// (|err| <= 3e-7): 6 FMA + 4 squarings + ONE MUFU (rcp).  p^16 overflows to +inf for z > ~9 -> erf = 1.
__device__ __forceinline__ float erf_pos_rcp16(float z) {
  float p = fmaf(0.0000430638f, z, 0.0002765672f);
  p = fmaf(p, z, 0.0001520143f);
  p = fmaf(p, z, 0.0092705272f);
  p = fmaf(p, z, 0.0422820123f);
  p = fmaf(p, z, 0.0705230784f);
  p = fmaf(p, z, 1.0f);
  p *= p;
  p *= p;
  p *= p;
  p *= p;
  return 1.0f - __fdividef(1.0f, p);
}
// log1p(u) for u in [0, 1] = u * q(u), degree-7 minimax-like fit (rel err 3.6e-7): no MUFU
__device__ __forceinline__ float log1p_unit(float u) {
  float q = fmaf(-0.0085746762f, u, 0.0442141923f);
  q = fmaf(q, u, -0.107853679f);
  q = fmaf(q, u, 0.17757024f);
  q = fmaf(q, u, -0.244996117f);
  q = fmaf(q, u, 0.332761766f);
  q = fmaf(q, u, -0.499974494f);
  q = fmaf(q, u, 0.99999981f);
  return q * u;
}
// SERF with 2 MUFU ops (ex2, rcp): softplus(x) = max(x,0) + log1p(exp(-|x|)), erf by 7.1.28.  |abs err| < 2e-5.
__device__ __forceinline__ float serf_fast(float x) {
  const float e = __expf(-fabsf(x));
  const float sp = fmaxf(x, 0.0f) + log1p_unit(e);
  return x * erf_pos_rcp16(sp);
}
// SERF' with 4 MUFU ops (ex2, rcp, rcp, ex2).  |abs err| < 3e-6.
__device__ __forceinline__ float dserf_fast(float x) {
  const float e = __expf(-fabsf(x));
  const float sp = fmaxf(x, 0.0f) + log1p_unit(e);
  const float er = erf_pos_rcp16(sp);
  const float E = __expf(-sp * sp);
  const float inv = __fdividef(1.0f, 1.0f + e);
  const float sig = x >= 0.0f ? inv : e * inv;
  // beyond the reference's clamp (x > 50) E underflows to 0 and the slope is erf(sp) = 1, as in serf.py
  return fmaf(x * 1.1283791670955126f * E, sig, er);
}
// SERF and SERF' together (shared softplus / erf): 4 MUFU for both
__device__ __forceinline__ void serf_both_fast(float x, float& a, float& d) {
  const float e = __expf(-fabsf(x));
  const float sp = fmaxf(x, 0.0f) + log1p_unit(e);
  const float er = erf_pos_rcp16(sp);
  const float E = __expf(-sp * sp);
  const float inv = __fdividef(1.0f, 1.0f + e);
  const float sig = x >= 0.0f ? inv : e * inv;
  a = x * er;
  d = fmaf(x * 1.1283791670955126f * E, sig, er);
}
__device__ __forceinline__ float gelu_fast(float x) {
  const ErfPos r = erf_pos_fast(fabsf(x) * 0.7071067811865476f);
  return 0.5f * x * (1.0f + copysignf(r.erf, x));
}
__device__ __forceinline__ float dgelu_fast(float x) {
  const ErfPos r = erf_pos_fast(fabsf(x) * 0.7071067811865476f);
  return fmaf(x * 0.3989422804014327f, r.E, 0.5f * (1.0f + copysignf(r.erf, x)));
}
template <int ACT> __device__ __forceinline__ float act_fast(float x) {
  if (ACT == MMVQA_ACT_SERF) return serf_fast(x);
  if (ACT == MMVQA_ACT_GELU) return gelu_fast(x);
  if (ACT == MMVQA_ACT_RELU) return fmaxf(x, 0.0f);
  return x;
}
template <int ACT> __device__ __forceinline__ void act_both_fast(float x, float& a, float& d) {
  if (ACT == MMVQA_ACT_SERF) {
    serf_both_fast(x, a, d);
  } else {
    a = act_fast<ACT>(x);
    if (ACT == MMVQA_ACT_GELU) d = dgelu_fast(x);
    else if (ACT == MMVQA_ACT_RELU) d = x > 0.0f ? 1.0f : 0.0f;
    else d = 1.0f;
  }
}
template <int ACT> __device__ __forceinline__ float dact_fast(float x) {
  if (ACT == MMVQA_ACT_SERF) return dserf_fast(x);
  if (ACT == MMVQA_ACT_GELU) return dgelu_fast(x);
  if (ACT == MMVQA_ACT_RELU) return x > 0.0f ? 1.0f : 0.0f;
  return 1.0f;
}

// ---------------------------------------------------------------------------------
// SERF through a shared-memory table (bf16 tensor-core epilogues with very many activations per output, i.e. the
// visual-token projector).  g(x) = erf(softplus(x)) is smooth and saturates (g = 0 below x = -18 to 2e-8, g = 1 above
// x = 6 to fp32 precision), so a cubic Hermite interpolant over 256 intervals of width 3/32 reproduces g to ~1e-6
// and g' to ~1e-5 with ONE 128-bit shared-memory load and ~20 FMA-pipe instructions per element, no MUFU:
//   serf(x) = x g(x),   serf'(x) = g(x) + x g'(x).
// Entry i holds the cubic of interval i in the local variable t = (x - x_i) / h:  g = c0 + t (c1 + t (c2 + t c3)) with
// c0 = g_i, c1 = h g'_i, c2 = 3 (g_{i+1} - g_i) - 2 h g'_i - h g'_{i+1}, c3 = h g'_i + h g'_{i+1} - 2 (g_{i+1} - g_i)
// (the Hermite interpolant, expanded once at fill time so the lookup is three FMAs); the nodes are evaluated with the accurate libdevice functions by
// the CTA itself in its prologue (no host-side state, valid inside CUDA-graph capture).
// The table is stored 8 times, replica k in the k-th 16-byte bank group: entry (i, k) sits at float4 index 8 i + k and
// lane l reads replica l & 7, so the 8 lanes of every quarter-warp phase of the 128-bit load hit 8 different bank
// groups whatever rows they ask for -- the lookup is bank-conflict free for arbitrary data (measured before:
// 9.6 wavefronts per load with one copy, the shared-memory pipe was the bound of the projector kernel).
// ---------------------------------------------------------------------------------
constexpr int SERF_TAB_N = 256;
// replicas: 8 = conflict free (32 KB); 4 halves the footprint where shared memory is tight (at most 2-way conflicts)
constexpr float SERF_TAB_X0 = -18.0f;
constexpr float SERF_TAB_H = 0.09375f;            // 24 / 256
constexpr float SERF_TAB_INV_H = 1.0f / 0.09375f;

__device__ __forceinline__ void serf_node(int i, float& g, float& m) {
  if (i <= 0) { g = 0.0f; m = 0.0f; return; }                 // flat tail: serf(x) = 0 (true value |x| 2e-8) below -18
  if (i >= SERF_TAB_N) { g = 1.0f; m = 0.0f; return; }        // erf(softplus(6)) = 1 in fp32: slope exactly 1 above
  const float x = SERF_TAB_X0 + SERF_TAB_H * (float)i;
  const float sp = log1pf(expf(x));
  g = erff(sp);
  m = SERF_TAB_H * 1.1283791670955126f * expf(-sp * sp) / (1.0f + expf(-x));
}
// fills the replicated table (shared memory, SERF_TAB_N * REP * 16 bytes); the caller synchronises afterwards
template <int REP>
__device__ __forceinline__ void serf_table_fill(float4* tab, int tid, int nthreads) {
  for (int i = tid; i < SERF_TAB_N; i += nthreads) {
    float g0, m0, g1, m1;
    serf_node(i, g0, m0);
    serf_node(i + 1, g1, m1);
    const float dl = g1 - g0;
    const float4 e = make_float4(g0, m0, 3.0f * dl - 2.0f * m0 - m1, m0 + m1 - 2.0f * dl);
#pragma unroll
    for (int k = 0; k < REP; ++k) tab[i * REP + k] = e;
  }
}
// table entry of x and the local variable t in [0, 1).  The interval index is taken from the mantissa of
// u + 1.5 * 2^23 added with round-down (floor(u) sits in the low bits, u - floor(u) follows with two FADDs): no F2I / I2F,
// which issue on the quarter-rate XU pipe, and a single FFMA + two FMNMX for the clamped coordinate.  The projector
// evaluates ~205 M activations per step and is bound by instruction issue (profiles/r01_ncu_vistok_extract.txt).
// `tab` here is this lane's handle from serf_tab_handle(): the shared-memory address of its replica minus the bias of
// the magic constant, so that the entry address is ONE multiply-add of the raw bits of u + 1.5 * 2^23.
template <int REP>
__device__ __forceinline__ uint32_t serf_tab_handle(const float4* tab) {
  return (uint32_t)__cvta_generic_to_shared(tab) + (threadIdx.x & (REP - 1)) * 16u - 0x4B400000u * (16u * REP);
}
template <int REP>
__device__ __forceinline__ float4 serf_tab_entry(uint32_t tab, float x, float& t) {
  const float u = fminf(fmaxf(fmaf(x, SERF_TAB_INV_H, -SERF_TAB_X0 * SERF_TAB_INV_H), 0.0f), (float)SERF_TAB_N - 0.001f);
  const float um = __fadd_rd(u, 12582912.0f);
  t = u - (um - 12582912.0f);
  float4 e;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
      : "=f"(e.x), "=f"(e.y), "=f"(e.z), "=f"(e.w)
      : "r"((uint32_t)__float_as_int(um) * (16u * REP) + tab));
  return e;
}
// a = serf(x), d = serf'(x)
template <int REP>
__device__ __forceinline__ void serf_both_tab(uint32_t tab, float x, float& a, float& d) {
  float t;
  const float4 e = serf_tab_entry<REP>(tab, x, t);
  const float q = fmaf(t, e.w, e.z);                 // c2 + t c3
  const float g = fmaf(t, fmaf(t, q, e.y), e.x);
  const float gp = fmaf(t, q + fmaf(t, e.w, q), e.y);   // c1 + t (2 c2 + 3 t c3)
  a = x * g;
  d = fmaf(x * SERF_TAB_INV_H, gp, g);
}
template <int REP>
__device__ __forceinline__ float serf_tab(uint32_t tab, float x) {
  float t;
  const float4 e = serf_tab_entry<REP>(tab, x, t);
  return x * fmaf(t, fmaf(t, fmaf(t, e.w, e.z), e.y), e.x);
}
template <int REP>
__device__ __forceinline__ float dserf_tab(uint32_t tab, float x) {
  float a, d;
  serf_both_tab<REP>(tab, x, a, d);
  return d;
}

// ---------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.0f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

// effective dropout seed: host seed + device step counter (NULL = eager mode, the host draws a new seed per call)
__device__ __forceinline__ unsigned long long seed_eff(unsigned long long seed, const unsigned long long* ctr) {
  return ctr ? seed + (*ctr) * 0x9E3779B97F4A7C15ull : seed;
}
// counter-based dropout keep decision (one 32-bit hash per element); identical in fwd and bwd
__device__ __forceinline__ uint32_t hash32(uint64_t seed, uint64_t idx) {
  uint64_t z = idx + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 16);
}

}  // namespace mmvqa
