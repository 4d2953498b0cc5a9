// gemm_tc_kernel.cuh -- the tcgen05 / TMEM / TMA GEMM kernel template and its launcher (see gemm_tc.cu for the
// design notes).  Instantiated per tile width in gemm_tc_bn*.cu so the translation units compile in parallel.
#pragma once
#include "gemm_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace mmvqa {

int tc_make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int nbatch,
                int64_t batch_rows, int box_inner, int box_rows, const char* what);
// 4-D map {64, rows, inner / 64, batch} over a stored [rows, inner] matrix: dimension 2 walks the 64-element chunks of
// the contiguous dimension, so ONE box {64, box_rows, box_chunks} brings several chunks (K-major operand: several
// k-blocks; MN-major operand: the 64-wide groups of a tile).  Needs inner % 64 == 0.
int tc_make_map_chunked(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int nbatch,
                        int64_t batch_rows, int box_rows, int box_chunks, const char* what);


constexpr int TC_BM = 128;      // UMMA M (cta_group::1)
constexpr int TC_BK = 64;       // 64 bf16 = one 128-byte swizzle row
constexpr int TC_UK = 16;       // UMMA K for 16-bit inputs
constexpr int TC_THREADS = 320;   // TMA warp + MMA warp + 8 epilogue warps

// ---------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_g(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait for an EARLIER tcgen05.ld into r[]: the registers are in/out operands, so no use of them can move above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
// UMMA shared-memory matrix descriptor, 128-byte swizzle (cute::UMMA::SmemDescriptor, version 1):
//  [0,14) start >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//  [46,48) version = 1 | [61,64) layout = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_TRACE(slot)                                                                              \
  do {                                                                                              \
    if (p.trace) p.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (slot)] = gtime(); \
  } while (0)

// KPS = 64-wide k-blocks per ring stage.  Measured (tools/ubench/stream_bench*.cu, the phase tracer): a ring stage costs
// a CTA ~0.3-0.4 us whatever its size up to 64 KB and whatever the ring depth, so small-K problems take FEWER, FATTER
// stages: one TMA box per operand brings KPS k-blocks (4-D tensor map, tc_make_map_chunked).
// SERF in the FF1 / FF2-dgrad epilogues (realformer.py:22-26 and its backward) from the shared-memory Hermite table of
// common.cuh instead of 3-4 MUFU per element: two replicas (8 KB) fit next to a 96 KB ring at two CTAs per SM.  The
// table is filled in the prologue, i.e. under the previous kernel's tail (programmatic dependent launch).
constexpr int TC_SERF_REP = 2;
constexpr int TC_SERF_TAB_BYTES = SERF_TAB_N * TC_SERF_REP * 16;

template <int BN, int STAGES_, int KPS_ = 1>
struct TcCfg {
  static constexpr int STAGES = STAGES_;
  static constexpr int KPS = KPS_;
  static constexpr int A_BYTES = TC_BM * TC_BK * 2;   // 16 KB per k-block
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = KPS * (A_BYTES + B_BYTES);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

// ---------------------------------------------------------------------------------
// epilogue: one thread = one accumulator row (TMEM lane); 16 columns per tcgen05.ld
// ---------------------------------------------------------------------------------
__device__ __forceinline__ bool al16(const void* base, int64_t elem_off, int elem_bytes) {
  return ((reinterpret_cast<uintptr_t>(base) + (uintptr_t)(elem_off * elem_bytes)) & 15) == 0;
}
__device__ __forceinline__ void load16_bf16(const __nv_bfloat16* src, int nvalid, float* out) {
  if (nvalid == 16 && al16(src, 0, 2)) {
    Vec16<__nv_bfloat16> a, b;
    a.load(src);
    b.load(src + 8);
#pragma unroll
    for (int j = 0; j < 8; ++j) { out[j] = a.get(j); out[8 + j] = b.get(j); }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) out[j] = j < nvalid ? __bfloat162float(src[j]) : 0.0f;
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, int nvalid, const float* v) {
  if (nvalid == 16 && al16(dst, 0, 2)) {
    reinterpret_cast<uint4*>(dst)[0] = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    reinterpret_cast<uint4*>(dst)[1] = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dst[j] = __float2bfloat16_rn(v[j]);
  }
}
__device__ __forceinline__ void store16_f32(float* dst, int nvalid, const float* v) {
  if (nvalid == 16 && al16(dst, 0, 4)) {
#pragma unroll
    for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dst[j] = v[j];
  }
}

// chunk `ci` (16 bf16 = two 128-bit registers) of a prefetched row slice -> 16 floats; the chunk index is
// resolved with selects so the array stays in registers
template <int NV>
__device__ __forceinline__ void unpack_pre(const uint4* pre, int ci, float* out) {
  uint4 lo = pre[0], hi = pre[NV > 1 ? 1 : 0];
#pragma unroll
  for (int q = 1; q < NV / 2; ++q) {
    if (ci == q) { lo = pre[2 * q]; hi = pre[2 * q + 1]; }
  }
  const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    out[2 * j] = __uint_as_float(w[j] << 16);
    out[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u);
  }
}

// ---------------------------------------------------------------------------------
// Coalesced output: "thread = accumulator row" means a direct store touches 32 different rows per instruction (32-byte
// pieces of 32 lines; measured: 7.9 us to write one 128 x 256 bf16 tile, more than its 4.5 us main loop).  Instead every
// epilogue warp parks 128-byte row pieces of its 32 rows in the ring's shared memory (free once the accumulator is
// complete; pitch 144 B -> conflict-free 128-bit accesses both ways) and writes them back 4 rows x 128 B per instruction.
// The pieces of a warp are written and read by that warp only: __syncwarp is the only synchronisation.
// ---------------------------------------------------------------------------------
constexpr int STG_PITCH = 144;                     // bytes per staged row piece (128 + 16 padding)
constexpr int STG_WARP_BYTES = 32 * STG_PITCH;     // 4608
constexpr int STG_BYTES = 8 * STG_WARP_BYTES;      // 36864 for the 8 epilogue warps (x2 with a second output)

__device__ __forceinline__ void stg_put16(uint8_t* stg, int lane, int unit, const float* v, bool as_bf16) {
  uint8_t* dst = stg + lane * STG_PITCH + unit * 16;
  if (as_bf16) {
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    *reinterpret_cast<uint4*>(dst + 16) = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]), pack_bf16x2(v[14], v[15]));
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<float4*>(dst + 16 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
}
// rows [0, rows_valid) x `units` 16-byte units of the warp's staged pieces -> global rows base + r * row_bytes
__device__ __forceinline__ void stg_flush(const uint8_t* stg, uint8_t* base, int64_t row_bytes, int rows_valid, int units, int lane) {
  __syncwarp();
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int idx = it * 32 + lane, r = idx >> 3, u = idx & 7;
    if (r < rows_valid && u < units)
      *reinterpret_cast<uint4*>(base + (int64_t)r * row_bytes + u * 16) = *reinterpret_cast<const uint4*>(stg + r * STG_PITCH + u * 16);
  }
  __syncwarp();
}

template <int EPI, int ACT, int BN>
__device__ __forceinline__ void tc_epilogue(const EpiParams& p, uint32_t tmem_row, int m, int n0, int bz, bool first,
                                            bool row_ok, bool has_acc, int c_begin, int64_t c_split_off,
                                            uint32_t tmem_full_bar, uint8_t* stg = nullptr, uint8_t* stg_aux = nullptr,
                                            int rows_valid = 0) {
  constexpr int NCH = BN / 2 / 16;                 // 16-column chunks this warp owns
  constexpr bool PRE = (EPI == MMVQA_EPI_RESIDUAL || EPI == MMVQA_EPI_DACT) && BN <= 128;
  const int lane = threadIdx.x & 31;
  float rowsum = 0.0f;
  float rscale = 0.0f;
  if ((EPI == MMVQA_EPI_DACT_SCALE || (EPI == MMVQA_EPI_STORE && p.rowscale != nullptr)) && row_ok)
    rscale = __ldg(p.rowscale + (int64_t)bz * p.M + m) * p.scale;
  const bool use_bias = p.bias != nullptr && first;
  const uint32_t thr = (uint32_t)(p.dropout_p * 4294967296.0);
  const float inv_keep = p.dropout_p > 0.0f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  // ---- everything that does not depend on the accumulator is fetched NOW, while the main loop is still running:
  // this thread's slice of the residual / pre-activation row (128-bit loads) and the bias (one column per lane,
  // broadcast with shuffles later).  The epilogue then starts with its operands in registers.
  uint4 pre[PRE ? NCH * 2 : 1];
  bool pre_ok = false;
  if (PRE) {
    const __nv_bfloat16* ax = reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + (int64_t)m * p.ld_aux_in + n0 + c_begin;
    pre_ok = row_ok && (n0 + c_begin + BN / 2 <= p.N) && al16(ax, 0, 2);
    if (pre_ok) {
#pragma unroll
      for (int i = 0; i < NCH * 2; ++i) pre[i] = __ldg(reinterpret_cast<const uint4*>(ax) + i);
    }
  }
  // (the bias is read per chunk with warp-uniform 128-bit loads: the ncu capture of the 4096 x 3072 x 768 GEMM showed the
  // epilogue at ~550 instructions per 16-column chunk, the kernel bound by issue slots and instruction fetch)
  const bool bias_vec = use_bias && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  if (threadIdx.x == 64) TC_TRACE(6);
  const unsigned long long dseed = (EPI == MMVQA_EPI_RESIDUAL && p.dropout_p > 0.0f) ? seed_eff(p.dropout_seed, p.seed_ctr) : 0ull;
  // the TMEM read of chunk ci + 1 is in flight while chunk ci is processed (a tcgen05.ld round trip is ~0.2 us)
  uint32_t rn[16];
  __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
  tmem_ld16(tmem_row + (uint32_t)c_begin, rn);
#pragma unroll 1
  for (int ci = 0; ci < NCH; ++ci) {
    const int c = c_begin + ci * 16;
    uint32_t r[16];
    __syncwarp();
    tmem_ld_wait16(rn);
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = rn[j];
    if (ci + 1 < NCH) tmem_ld16(tmem_row + (uint32_t)(c + 16), rn);
    if (threadIdx.x == 64 && ci == 0) TC_TRACE(9);
    const int nb = n0 + c;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = has_acc ? __uint_as_float(r[j]) : 0.0f;
    if (use_bias) {
      if (bias_vec && nb + 16 <= p.N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j);
          v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
      }
    }
    if (row_ok && nb < p.N) {
      const int nvalid = min(16, p.N - nb);
      const int64_t coff = (int64_t)bz * p.c_batch_stride + c_split_off + (int64_t)m * p.ldc + nb;
      if (EPI == MMVQA_EPI_ACT_ROWSUM) {
        if (p.aux_out) {
          float dv_[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float a_;
            act_both_fast<ACT>(v[j], a_, dv_[j]);
            if (j < nvalid) rowsum += a_;
          }
          store16_bf16(reinterpret_cast<__nv_bfloat16*>(p.aux_out) + ((int64_t)bz * p.M + m) * p.ld_aux_out + nb, nvalid, dv_);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nvalid) rowsum += act_fast<ACT>(v[j]);
        }
        goto next_chunk;
      }
      if (EPI == MMVQA_EPI_STORE && p.rowscale != nullptr) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= rscale;
      }
      if (EPI == MMVQA_EPI_ACT) {
        if (p.aux_out) {
          if (stg_aux) stg_put16(stg_aux, lane, (ci & 3) * 2, v, true);
          else store16_bf16(reinterpret_cast<__nv_bfloat16*>(p.aux_out) + (int64_t)m * p.ld_aux_out + nb, nvalid, v);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act_fast<ACT>(v[j]);
      } else if (EPI == MMVQA_EPI_RESIDUAL) {
        float a[16];
        if (PRE && pre_ok) unpack_pre<PRE ? NCH * 2 : 1>(pre, ci, a);
        else load16_bf16(reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + (int64_t)m * p.ld_aux_in + nb, nvalid, a);
        if (p.dropout_p > 0.0f) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            v[j] = hash32(dseed, (uint64_t)m * (uint64_t)p.N + (uint64_t)(nb + j)) >= thr ? v[j] * inv_keep : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += a[j];
      } else if (EPI == MMVQA_EPI_DACT) {
        float a[16];
        if (PRE && pre_ok) unpack_pre<PRE ? NCH * 2 : 1>(pre, ci, a);
        else load16_bf16(reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + (int64_t)m * p.ld_aux_in + nb, nvalid, a);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= dact_fast<ACT>(a[j]);
      } else if (EPI == MMVQA_EPI_DACT_SCALE) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = dact_fast<ACT>(v[j]) * rscale;
      }
      if (p.accumulate) {
        float* c32 = reinterpret_cast<float*>(p.C) + coff;
        if (nvalid == 16 && al16(c32, 0, 4)) {   // 128-bit vector reductions (red.global.add.v4.f32): 4x fewer L2 atomics
#pragma unroll
          for (int j = 0; j < 4; ++j)
            atomicAdd(reinterpret_cast<float4*>(c32) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nvalid) atomicAdd(c32 + j, v[j]);
        }
      } else if (stg) {
        stg_put16(stg, lane, p.c_bf16 ? (ci & 3) * 2 : (ci & 1) * 4, v, p.c_bf16 != 0);
      } else if (p.c_bf16) {
        store16_bf16(reinterpret_cast<__nv_bfloat16*>(p.C) + coff, nvalid, v);
      } else {
        store16_f32(reinterpret_cast<float*>(p.C) + coff, nvalid, v);
      }
      if (nvalid < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j >= nvalid) v[j] = 0.0f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.0f;   // rows / columns outside the problem add nothing to the column sums
    }
    if (threadIdx.x == 64 && c == c_begin) TC_TRACE(10);
    if (EPI != MMVQA_EPI_ACT_ROWSUM && p.colsum_out != nullptr && nb < p.N) {
      // column sums over the warp's 32 rows: transposing butterfly, 16 shuffles; lane l ends with column
      // 8*b4 + 4*b3 + 2*b2 + b1 of the chunk (b_k = bit k of l), duplicated on the lane pair (l, l^1)
      __syncwarp();
      {
        const bool hi = lane & 16;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float send = hi ? v[k] : v[k + 8];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
          v[k] = (hi ? v[k + 8] : v[k]) + recv;
        }
      }
      {
        const bool hi = lane & 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float send = hi ? v[k] : v[k + 4];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
          v[k] = (hi ? v[k + 4] : v[k]) + recv;
        }
      }
      {
        const bool hi = lane & 4;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float send = hi ? v[k] : v[k + 2];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
          v[k] = (hi ? v[k + 2] : v[k]) + recv;
        }
      }
      {
        const bool hi = lane & 2;
        const float send = hi ? v[0] : v[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        v[0] = (hi ? v[1] : v[0]) + recv;
      }
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
      const int col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
      if ((lane & 1) == 0 && nb + col < p.N) atomicAdd(p.colsum_out + nb + col, v[0]);
    }
  next_chunk:;
    if (EPI != MMVQA_EPI_ACT_ROWSUM && stg != nullptr) {   // warp-uniform: write the staged pass back, 4 rows x 128 B per instruction
      const int per_pass = p.c_bf16 ? 4 : 2;
      if ((ci % per_pass) == per_pass - 1 || ci == NCH - 1) {
        const int ci0 = ci - (ci % per_pass);
        const int col0 = n0 + c_begin + ci0 * 16;
        if (col0 < p.N) {
          const int ncols = min(p.N - col0, (ci - ci0 + 1) * 16);
          const int elem = p.c_bf16 ? 2 : 4;
          uint8_t* base = reinterpret_cast<uint8_t*>(p.C) +
                          ((int64_t)bz * p.c_batch_stride + c_split_off + (int64_t)(m - lane) * p.ldc + col0) * elem;
          stg_flush(stg, base, p.ldc * elem, rows_valid, ncols * elem / 16, lane);
          if (EPI == MMVQA_EPI_ACT && stg_aux != nullptr) {
            uint8_t* abase = reinterpret_cast<uint8_t*>(p.aux_out) + ((int64_t)(m - lane) * p.ld_aux_out + col0) * 2;
            stg_flush(stg_aux, abase, p.ld_aux_out * 2, rows_valid, ncols * 2 / 16, lane);
          }
        }
      }
    }
  }
  if (EPI == MMVQA_EPI_ACT_ROWSUM && row_ok) atomicAdd(p.rowsum_out + (int64_t)bz * p.M + m, rowsum * p.scale);
}

// ---------------------------------------------------------------------------------
// Lean epilogue for the common case: the warp's 32 rows x BN/2 columns lie inside the problem, the output is staged
// through shared memory, no column sums / row scales / atomics.  The general tc_epilogue above pays ~190 instructions and
// a dozen taken branches per 16-column chunk for its guards (ncu source page, profiles/r02_gemm_epilogue.txt: the
// 4096 x 3072 x 768 GEMM spent 4.2 us per tile in it against a 4.9 us main loop); here a chunk is one TMEM read, the
// epilogue arithmetic, two 128-bit shared-memory stores, and every 128 bytes of row a write-back of 4 rows x 128 B per
// instruction.  The TMEM read and the residual / pre-activation row of chunk i + 1 are in flight while chunk i is
// processed.  Same arithmetic, same dropout index, same rounding as tc_epilogue.
// ---------------------------------------------------------------------------------
template <int EPI, int ACT, int BN>
__device__ __forceinline__ void tc_epilogue_fast(const EpiParams& p, uint32_t tmem_row, int m, int n0, int bz, bool first,
                                                 int c_begin, int64_t c_split_off, uint32_t tmem_full_bar, uint8_t* stg,
                                                 uint8_t* stg_aux, uint32_t tab = 0) {
  constexpr int NCH = BN / 2 / 16;
  constexpr bool AUXIN = (EPI == MMVQA_EPI_RESIDUAL || EPI == MMVQA_EPI_DACT);
  const int lane = threadIdx.x & 31;
  const int col0 = n0 + c_begin;
  const bool use_bias = p.bias != nullptr && first;
  const uint4* ax = nullptr;
  uint4 a_lo = make_uint4(0, 0, 0, 0), a_hi = a_lo;
  if (AUXIN) {   // first chunk of the residual / pre-activation row: fetched while the main loop is still running
    ax = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.aux_in) + (int64_t)m * p.ld_aux_in + col0);
    a_lo = __ldg(ax);
    a_hi = __ldg(ax + 1);
  }
  const uint32_t thr = (uint32_t)(p.dropout_p * 4294967296.0);
  const float inv_keep = p.dropout_p > 0.0f ? 1.0f / (1.0f - p.dropout_p) : 1.0f;
  const bool c_bf16 = p.c_bf16 != 0;
  const int elem = c_bf16 ? 2 : 4;
  const int per_pass = c_bf16 ? 4 : 2;                       // chunks per 128 bytes of output row
  const int units = (NCH < per_pass ? NCH : per_pass) * (c_bf16 ? 2 : 4);
  // write-back geometry: lane -> row (lane >> 3) + 4 it of the warp's 32, 16-byte unit lane & 7
  const int64_t row_bytes = p.ldc * elem;
  uint8_t* cbase = reinterpret_cast<uint8_t*>(p.C) +
                   ((int64_t)bz * p.c_batch_stride + c_split_off + (int64_t)(m - lane + (lane >> 3)) * p.ldc + col0) * elem + (lane & 7) * 16;
  const uint32_t soff = (lane >> 3) * STG_PITCH + (lane & 7) * 16;
  uint8_t* abase = nullptr;
  int64_t arow_bytes = 0;
  if (EPI == MMVQA_EPI_ACT && stg_aux != nullptr) {
    arow_bytes = p.ld_aux_out * 2;
    abase = reinterpret_cast<uint8_t*>(p.aux_out) + ((int64_t)(m - lane + (lane >> 3)) * p.ld_aux_out + col0) * 2 + (lane & 7) * 16;
  }
  mbar_wait(tmem_full_bar, 0);
  tc_fence_after();
  if (threadIdx.x == 64) TC_TRACE(6);
  const unsigned long long dseed = (EPI == MMVQA_EPI_RESIDUAL && p.dropout_p > 0.0f) ? seed_eff(p.dropout_seed, p.seed_ctr) : 0ull;
  uint32_t rn[16];
  __syncwarp();  // tcgen05.ld is .sync.aligned: the warp must be converged
  tmem_ld16(tmem_row + (uint32_t)c_begin, rn);
#pragma unroll 1
  for (int ci = 0; ci < NCH; ++ci) {
    uint32_t r[16];
    tmem_ld_wait16(rn);
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = rn[j];
    const uint4 q_lo = a_lo, q_hi = a_hi;
    if (ci + 1 < NCH) {
      tmem_ld16(tmem_row + (uint32_t)(c_begin + (ci + 1) * 16), rn);
      if (AUXIN) {
        a_lo = __ldg(ax + 2 * (ci + 1));
        a_hi = __ldg(ax + 2 * (ci + 1) + 1);
      }
    }
    const int nb = col0 + ci * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    if (use_bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j);
        v[4 * j] += b4.x; v[4 * j + 1] += b4.y; v[4 * j + 2] += b4.z; v[4 * j + 3] += b4.w;
      }
    }
    if (EPI == MMVQA_EPI_ACT) {
      if (stg_aux != nullptr) stg_put16(stg_aux, lane, (ci & 3) * 2, v, true);
      if (ACT == MMVQA_ACT_SERF && tab != 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = serf_tab<TC_SERF_REP>(tab, v[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = act_fast<ACT>(v[j]);
      }
    } else if (AUXIN) {
      const uint32_t w[8] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w, q_hi.x, q_hi.y, q_hi.z, q_hi.w};
      if (EPI == MMVQA_EPI_RESIDUAL) {
        if (p.dropout_p > 0.0f) {
          const uint64_t i0 = (uint64_t)m * (uint64_t)p.N + (uint64_t)nb;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = hash32(dseed, i0 + (uint64_t)j) >= thr ? v[j] * inv_keep : 0.0f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[2 * j] += __uint_as_float(w[j] << 16);
          v[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
        }
      } else if (ACT == MMVQA_ACT_SERF && tab != 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[2 * j] *= dserf_tab<TC_SERF_REP>(tab, __uint_as_float(w[j] << 16));
          v[2 * j + 1] *= dserf_tab<TC_SERF_REP>(tab, __uint_as_float(w[j] & 0xffff0000u));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[2 * j] *= dact_fast<ACT>(__uint_as_float(w[j] << 16));
          v[2 * j + 1] *= dact_fast<ACT>(__uint_as_float(w[j] & 0xffff0000u));
        }
      }
    }
    stg_put16(stg, lane, c_bf16 ? (ci & 3) * 2 : (ci & 1) * 4, v, c_bf16);
    if (((ci + 1) & (per_pass - 1)) == 0 || ci == NCH - 1) {   // warp-uniform: 128 bytes of every row are staged
      const int ci0 = ci & ~(per_pass - 1);
      __syncwarp();
      if ((lane & 7) < units) {
        uint8_t* g = cbase + (int64_t)ci0 * 16 * elem;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          *reinterpret_cast<uint4*>(g + it * 4 * row_bytes) = *reinterpret_cast<const uint4*>(stg + soff + it * 4 * STG_PITCH);
        if (EPI == MMVQA_EPI_ACT && stg_aux != nullptr) {
          uint8_t* ga = abase + (int64_t)ci0 * 32;
#pragma unroll
          for (int it = 0; it < 8; ++it)
            *reinterpret_cast<uint4*>(ga + it * 4 * arow_bytes) = *reinterpret_cast<const uint4*>(stg_aux + soff + it * 4 * STG_PITCH);
        }
      }
      __syncwarp();
    }
  }
}

template <int EPI, int BN>
__device__ __forceinline__ void tc_epilogue_fast_act(const EpiParams& p, uint32_t tmem_row, int m, int n0, int bz, bool first,
                                                     int c_begin, uint32_t tmem_full_bar, uint8_t* stg, uint8_t* stg_aux,
                                                     uint32_t tab = 0) {
  switch (p.act) {
    case MMVQA_ACT_SERF: tc_epilogue_fast<EPI, MMVQA_ACT_SERF, BN>(p, tmem_row, m, n0, bz, first, c_begin, 0, tmem_full_bar, stg, stg_aux, tab); break;
    case MMVQA_ACT_GELU: tc_epilogue_fast<EPI, MMVQA_ACT_GELU, BN>(p, tmem_row, m, n0, bz, first, c_begin, 0, tmem_full_bar, stg, stg_aux); break;
    case MMVQA_ACT_RELU: tc_epilogue_fast<EPI, MMVQA_ACT_RELU, BN>(p, tmem_row, m, n0, bz, first, c_begin, 0, tmem_full_bar, stg, stg_aux); break;
    default: tc_epilogue_fast<EPI, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, c_begin, 0, tmem_full_bar, stg, stg_aux); break;
  }
}

template <int EPI, int BN>
__device__ __forceinline__ void tc_epilogue_act(const EpiParams& p, uint32_t tmem_row, int m, int n0, int bz, bool first,
                                                bool row_ok, bool has_acc, int c_begin, uint32_t tmem_full_bar,
                                                uint8_t* stg = nullptr, uint8_t* stg_aux = nullptr, int rows_valid = 0) {
  const int64_t c_split_off = 0;   // activation epilogues never run split-K
  switch (p.act) {
    case MMVQA_ACT_SERF: tc_epilogue<EPI, MMVQA_ACT_SERF, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, stg_aux, rows_valid); break;
    case MMVQA_ACT_GELU: tc_epilogue<EPI, MMVQA_ACT_GELU, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, stg_aux, rows_valid); break;
    case MMVQA_ACT_RELU: tc_epilogue<EPI, MMVQA_ACT_RELU, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, stg_aux, rows_valid); break;
    default: tc_epilogue<EPI, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, stg_aux, rows_valid); break;
  }
}

// ---------------------------------------------------------------------------------
// kernel: one 128 x BN output tile (of one batch entry / one K split) per CTA
// ---------------------------------------------------------------------------------
template <int BN, bool A_MN, bool B_MN, int STAGES, int KPS = 1>
__global__ void __launch_bounds__(TC_THREADS, 2) gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, EpiParams p,
                                                             int a_batched, int b_batched, int b_static) {
  using Cfg = TcCfg<BN, STAGES, KPS>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + Cfg::STAGES * Cfg::STAGE_BYTES;
  // barriers: full[s] at +8*s, empty[s] at +8*(STAGES+s), tmem_full at +8*2*STAGES, tmem ptr after
  const uint32_t tmem_full_bar = bar_base + 8 * 2 * Cfg::STAGES;
  const uint32_t tmem_ptr_addr = tmem_full_bar + 8;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 8 * 2 * Cfg::STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * TC_BM;
  const int bz = blockIdx.z / p.split_k, ks = blockIdx.z % p.split_k;
  const int kblocks = (p.K + TC_BK - 1) / TC_BK;
  const int kb_per = (kblocks + p.split_k - 1) / p.split_k;
  const int kb0 = ks * kb_per;
  const int kb1 = min(kblocks, kb0 + kb_per);
  const int nkb = max(0, kb1 - kb0);
  const int nst = (nkb + KPS - 1) / KPS;          // ring stages of this CTA (KPS k-blocks each)
  if (threadIdx.x == 0) TC_TRACE(0);

  // SERF table behind the barrier block (only when the launcher reserved the bytes: p.serf_tab)
  float4* serf_tab_ptr = reinterpret_cast<float4*>(smem_gen + Cfg::STAGES * Cfg::STAGE_BYTES + 256);
  if (p.serf_tab) serf_table_fill<TC_SERF_REP>(serf_tab_ptr, threadIdx.x, TC_THREADS);
  if (warp == 0) {
    tmem_alloc(tmem_ptr_addr, Cfg::TMEM_COLS);
  } else if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      mbar_init(bar_base + 8 * s, 1);
      mbar_init(bar_base + 8 * (Cfg::STAGES + s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_ptr_gen;
  // operand loads of ring stage i (k-blocks kb0 + i*KPS ...) into the slot at `sa` / `sb`.
  // KPS == 1: the 3-D maps {inner, rows, batch}; KPS > 1: the 4-D chunked maps {64, rows, chunk, batch} -- one box per
  // operand carries KPS k-blocks (K-major: box {64, rows, KPS}; MN-major: box {64, 64 KPS k-rows, tile / 64 groups},
  // so the 64-wide MN groups are 8 KB * KPS apart).  Out-of-range chunks / rows are zero-filled by the TMA unit.
  auto load_a = [&](uint32_t sa, uint32_t full, int i) {
    const int kb = kb0 + i * KPS;
    const int zb = a_batched ? bz : 0;
    if (KPS == 1) {
      if (A_MN) {  // stored [K, M]: two boxes of 64 (m) x 64 (k)
        tma_load_3d(sa, &tmA, full, m0, kb * TC_BK, zb);
        tma_load_3d(sa + 8192, &tmA, full, m0 + 64, kb * TC_BK, zb);
      } else {     // stored [M, K]: one box of 64 (k) x 128 (m)
        tma_load_3d(sa, &tmA, full, kb * TC_BK, m0, zb);
      }
    } else {
      if (A_MN) tma_load_4d_g(sa, &tmA, full, 0, kb * TC_BK, m0 / 64, zb);
      else tma_load_4d_g(sa, &tmA, full, 0, m0, kb, zb);
    }
  };
  auto load_b = [&](uint32_t sb, uint32_t full, int i) {
    const int kb = kb0 + i * KPS;
    const int zb = b_batched ? bz : 0;
    if (KPS == 1) {
      if (B_MN) {  // stored [K, N]: BN/64 boxes of 64 (n) x 64 (k)
#pragma unroll
        for (int j = 0; j < BN / 64; ++j) tma_load_3d(sb + j * 8192, &tmB, full, n0 + j * 64, kb * TC_BK, zb);
      } else {     // stored [N, K]: one box of 64 (k) x BN (n)
        tma_load_3d(sb, &tmB, full, kb * TC_BK, n0, zb);
      }
    } else {
      if (B_MN) tma_load_4d_g(sb, &tmB, full, 0, kb * TC_BK, n0 / 64, zb);
      else tma_load_4d_g(sb, &tmB, full, 0, n0, kb, zb);
    }
  };
  // Programmatic dependent launch: this CTA may be resident while the previous kernel on the stream is still running.
  // The weight operand (b_static) does not depend on that kernel, so its first ring slots are requested now; the
  // activation operand, the epilogue inputs and every store wait for griddepcontrol.wait below.
  int b_ahead = 0;
  if (threadIdx.x == 0) TC_TRACE(1);
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (b_static) {
      b_ahead = nst < Cfg::STAGES ? nst : Cfg::STAGES;
      for (int i = 0; i < b_ahead; ++i) {
        const uint32_t full = bar_base + 8 * i;
        mbar_expect_tx(full, Cfg::STAGE_BYTES);
        load_b(smem_base + i * Cfg::STAGE_BYTES + KPS * Cfg::A_BYTES, full, i);
      }
    }
  }
  pdl_wait();      // everything above overlapped the previous kernel; dependent global memory is touched below
  pdl_trigger();
  if (threadIdx.x == 0) TC_TRACE(2);

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int i = 0; i < nst; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
        const uint32_t full = bar_base + 8 * s;
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES, sb = sa + KPS * Cfg::A_BYTES;
        if (i < b_ahead) {   // slot armed and its B tile already in flight: only A is missing
          load_a(sa, full, i);
          continue;
        }
        mbar_wait(bar_base + 8 * (Cfg::STAGES + s), ph ^ 1u);
        mbar_expect_tx(full, Cfg::STAGE_BYTES);
        load_a(sa, full, i);
        load_b(sb, full, i);
      }
      TC_TRACE(3);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int i = 0; i < nst; ++i) {
        const int s = i % Cfg::STAGES;
        const uint32_t ph = (uint32_t)(i / Cfg::STAGES) & 1u;
        mbar_wait(bar_base + 8 * s, ph);
        tc_fence_after();
        if (i == 0) TC_TRACE(4);
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES, sb = sa + KPS * Cfg::A_BYTES;
#pragma unroll
        for (int c = 0; c < KPS; ++c) {
          if (i * KPS + c < nkb) {   // k-blocks beyond this CTA's range (K tail / next split) were loaded but are not used
#pragma unroll
            for (int j = 0; j < TC_BK / TC_UK; ++j) {
              // K-major: 16 bf16 = 32 bytes inside the swizzled 128-byte row; SBO = 8 rows * 128 B; k-block c is one
              //   [rows][128 B] tile further.
              // MN-major: 16 k-rows = 2 groups of 8 rows (1024 B each); LBO = next 64-wide MN group (8 KB * KPS away);
              //   k-block c is 64 k-rows = 8 KB further inside every group.
              const uint64_t ad = A_MN ? make_sdesc(sa + c * 8192 + j * 2048, 8192 * KPS, 1024)
                                       : make_sdesc(sa + c * Cfg::A_BYTES + j * 32, 16, 1024);
              const uint64_t bd = B_MN ? make_sdesc(sb + c * 8192 + j * 2048, 8192 * KPS, 1024)
                                       : make_sdesc(sb + c * Cfg::B_BYTES + j * 32, 16, 1024);
              umma_bf16(tmem_acc, ad, bd, idesc, (i > 0 || c > 0 || j > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(bar_base + 8 * (Cfg::STAGES + s));  // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full_bar);                       // accumulator complete
      TC_TRACE(5);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global.  8 warps: two per TMEM lane group, each takes half the columns =====
    const int g = warp & 3;                 // TMEM lane group this warp may read
    const int c_begin = ((warp - 2) >> 2) * (BN / 2);
    const int m = m0 + g * 32 + lane;
    const bool first = (ks == 0);
    // (the wait for the accumulator barrier is inside tc_epilogue, after its operand prefetch)
    const bool row_ok = (m < p.M) && !(nkb == 0 && ks != 0 && p.c_split_stride == 0);
    const int64_t c_split_off = (int64_t)ks * p.c_split_stride;
    const uint32_t tmem_row = tmem_acc + ((uint32_t)(g * 32) << 16);
    const bool has_acc = nkb > 0;
    // staged, coalesced output through the ring's shared memory (free once the accumulator barrier has completed)
    const int elem = p.c_bf16 ? 2 : 4;
    const bool stage_ok = p.C != nullptr && !p.accumulate && p.epilogue != MMVQA_EPI_ACT_ROWSUM && !p.no_stage &&
                          (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (p.ldc * elem) % 16 == 0 && ((int64_t)p.N * elem) % 16 == 0 &&
                          (p.c_batch_stride * elem) % 16 == 0 && (p.c_split_stride * elem) % 16 == 0;
    const bool aux_ok = stage_ok && p.epilogue == MMVQA_EPI_ACT && p.aux_out != nullptr &&
                        Cfg::STAGES * Cfg::STAGE_BYTES >= 2 * STG_BYTES && (reinterpret_cast<uintptr_t>(p.aux_out) & 15) == 0 &&
                        (p.ld_aux_out * 2) % 16 == 0;
    uint8_t* stg = stage_ok ? smem_gen + (warp - 2) * STG_WARP_BYTES : nullptr;
    uint8_t* stg_aux = aux_ok ? smem_gen + STG_BYTES + (warp - 2) * STG_WARP_BYTES : nullptr;
    const bool cta_ok = !(nkb == 0 && ks != 0 && p.c_split_stride == 0);
    const int rows_valid = cta_ok ? max(0, min(32, p.M - (m0 + g * 32))) : 0;
    // lean path: every row and column of this warp's 32 x BN/2 block is inside the problem and nothing needs a guard
    const bool aux_in_al = (reinterpret_cast<uintptr_t>(p.aux_in) & 15) == 0 && (p.ld_aux_in * 2) % 16 == 0;
    bool fast_ok = stage_ok && has_acc && rows_valid == 32 && n0 + c_begin + BN / 2 <= p.N && p.colsum_out == nullptr &&
                   (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    switch (p.epilogue) {
      case MMVQA_EPI_STORE: fast_ok = fast_ok && p.rowscale == nullptr; break;
      case MMVQA_EPI_ACT: fast_ok = fast_ok && (p.aux_out == nullptr || (aux_ok && p.c_bf16)); break;
      case MMVQA_EPI_RESIDUAL:
      case MMVQA_EPI_DACT: fast_ok = fast_ok && aux_in_al; break;
      default: fast_ok = false;
    }
    const uint32_t serf_h = p.serf_tab ? serf_tab_handle<TC_SERF_REP>(serf_tab_ptr) : 0u;
    if (fast_ok) {
      switch (p.epilogue) {
        case MMVQA_EPI_ACT: tc_epilogue_fast_act<MMVQA_EPI_ACT, BN>(p, tmem_row, m, n0, bz, first, c_begin, tmem_full_bar, stg, stg_aux, serf_h); break;
        case MMVQA_EPI_RESIDUAL: tc_epilogue_fast<MMVQA_EPI_RESIDUAL, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, c_begin, c_split_off, tmem_full_bar, stg, nullptr); break;
        case MMVQA_EPI_DACT: tc_epilogue_fast_act<MMVQA_EPI_DACT, BN>(p, tmem_row, m, n0, bz, first, c_begin, tmem_full_bar, stg, nullptr, serf_h); break;
        default: tc_epilogue_fast<MMVQA_EPI_STORE, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, c_begin, c_split_off, tmem_full_bar, stg, nullptr); break;
      }
    } else
    switch (p.epilogue) {
      case MMVQA_EPI_ACT: tc_epilogue_act<MMVQA_EPI_ACT, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, tmem_full_bar, stg, stg_aux, rows_valid); break;
      case MMVQA_EPI_RESIDUAL: tc_epilogue<MMVQA_EPI_RESIDUAL, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, nullptr, rows_valid); break;
      case MMVQA_EPI_DACT: tc_epilogue_act<MMVQA_EPI_DACT, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, tmem_full_bar, stg, nullptr, rows_valid); break;
      case MMVQA_EPI_ACT_ROWSUM: tc_epilogue_act<MMVQA_EPI_ACT_ROWSUM, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, tmem_full_bar); break;
      case MMVQA_EPI_DACT_SCALE: tc_epilogue_act<MMVQA_EPI_DACT_SCALE, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, tmem_full_bar, stg, nullptr, rows_valid); break;
      default: tc_epilogue<MMVQA_EPI_STORE, MMVQA_ACT_NONE, BN>(p, tmem_row, m, n0, bz, first, row_ok, has_acc, c_begin, c_split_off, tmem_full_bar, stg, nullptr, rows_valid); break;
    }
    if (threadIdx.x == 64) TC_TRACE(7);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_TRACE(8);
  if (warp == 0) tmem_dealloc(tmem_acc, Cfg::TMEM_COLS);
}

// can this operand be addressed through the chunked map?  (the 64-element chunks must tile its contiguous dimension)
static inline bool tc_chunkable(const mmvqa_gemm_args* a) {
  const bool a_ok = a->a_trans ? (a->M % 64 == 0) : (a->K % 64 == 0);
  const bool b_ok = a->b_trans ? (a->N % 64 == 0) : (a->K % 64 == 0);
  return a_ok && b_ok;
}

template <int BN, bool A_MN, bool B_MN, int STAGES, int KPS = 1>
static int launch_tc(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  using Cfg = TcCfg<BN, STAGES, KPS>;
  CUtensorMap tmA, tmB;
  int rc;
  if (KPS == 1) {
    // A: K-major stored [M, K] -> inner K, box 64 x 128;  MN-major stored [K, M] -> inner M, box 64 x 64
    if (A_MN) rc = tc_make_map(&tmA, a->A, a->M, a->K, a->lda, a->batch, a->a_batch_rows, 64, 64, "A");
    else rc = tc_make_map(&tmA, a->A, a->K, a->M, a->lda, a->batch, a->a_batch_rows, 64, TC_BM, "A");
    if (rc) return rc;
    if (B_MN) rc = tc_make_map(&tmB, a->B, a->N, a->K, a->ldb, a->batch, a->b_batch_rows, 64, 64, "B");
    else rc = tc_make_map(&tmB, a->B, a->K, a->N, a->ldb, a->batch, a->b_batch_rows, 64, BN, "B");
    if (rc) return rc;
  } else {
    MMVQA_REQUIRE(tc_chunkable(a), "gemm(bf16): multi-k-block stages need contiguous dimensions that are multiples of 64");
    // K-major stored [rows, K]: box {64, tile rows, KPS chunks};  MN-major stored [K, MN]: box {64, 64 KPS k-rows, tile / 64}
    if (A_MN) rc = tc_make_map_chunked(&tmA, a->A, a->M, a->K, a->lda, a->batch, a->a_batch_rows, 64 * KPS, TC_BM / 64, "A");
    else rc = tc_make_map_chunked(&tmA, a->A, a->K, a->M, a->lda, a->batch, a->a_batch_rows, TC_BM, KPS, "A");
    if (rc) return rc;
    if (B_MN) rc = tc_make_map_chunked(&tmB, a->B, a->N, a->K, a->ldb, a->batch, a->b_batch_rows, 64 * KPS, BN / 64, "B");
    else rc = tc_make_map_chunked(&tmB, a->B, a->K, a->N, a->ldb, a->batch, a->b_batch_rows, BN, KPS, "B");
    if (rc) return rc;
  }
  auto kern = gemm_tc_kernel<BN, A_MN, B_MN, STAGES, KPS>;
  // EXPERIMENT (MMVQA_WGRAD_SOLO): weight-gradient GEMMs (both operands MN-major) ask for > half of the shared memory and
  // carry no programmatic edge, so at most one of their CTAs lives on an SM and the main chain keeps half of its registers
  static const int solo_env = getenv("MMVQA_WGRAD_SOLO") ? atoi(getenv("MMVQA_WGRAD_SOLO")) : 0;
  const bool solo = solo_env != 0 && A_MN && B_MN;
  constexpr int SOLO_SMEM = Cfg::SMEM > 116 * 1024 ? Cfg::SMEM : 116 * 1024;
  constexpr int TAB_SMEM = Cfg::SMEM + TC_SERF_TAB_BYTES;
  static_assert(TAB_SMEM <= 227 * 1024, "ring + SERF table exceed the shared memory of an SM");
  static bool attr_set = false;
  if (!attr_set) {
    MMVQA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (A_MN && B_MN) ? (SOLO_SMEM > TAB_SMEM ? SOLO_SMEM : TAB_SMEM) : TAB_SMEM));
    attr_set = true;
  }
  // SERF / SERF' epilogues of the encoder's feed-forward GEMMs read the activation from a shared-memory table
  // (opt-in, MMVQA_TC_SERF_TAB=1: measured at the flagship shape it is no faster than the MUFU formulas -- FF1 + SERF
  // 11.5 vs 11.0 us warm, FF2 dgrad 10.95 vs 11.2 us -- the epilogue's cost is its stores and instruction count, not MUFU)
  static const bool no_tab = getenv("MMVQA_TC_SERF_TAB") == nullptr;
  EpiParams epl = ep;
  epl.serf_tab = (!no_tab && a->act == MMVQA_ACT_SERF && (a->epilogue == MMVQA_EPI_ACT || a->epilogue == MMVQA_EPI_DACT)) ? 1 : 0;
  const size_t smem_bytes = epl.serf_tab ? (size_t)TAB_SMEM : (size_t)Cfg::SMEM;
  dim3 grid((a->N + BN - 1) / BN, (a->M + TC_BM - 1) / TC_BM, a->batch * a->split_k);
  MMVQA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm(bf16): grid too large");
  if (solo) {
    const size_t sm_bytes = (solo_env & 2) ? (size_t)Cfg::SMEM : (size_t)SOLO_SMEM;
    MMVQA_CUDA(launch_plain(kern, grid, dim3(TC_THREADS), sm_bytes, st, tmA, tmB, ep,
                            (a->batch > 1 && a->a_batch_rows > 0) ? 1 : 0, (a->batch > 1 && a->b_batch_rows > 0) ? 1 : 0, 0));
    MMVQA_LAUNCHED("gemm_tc_bf16");
    return MMVQA_OK;
  }
  MMVQA_CUDA(launch_pdl(kern, grid, dim3(TC_THREADS), smem_bytes, st, tmA, tmB, epl,
                        (a->batch > 1 && a->a_batch_rows > 0) ? 1 : 0, (a->batch > 1 && a->b_batch_rows > 0) ? 1 : 0,
                        (a->b_static && pdl_enabled()) ? 1 : 0));
  MMVQA_LAUNCHED("gemm_tc_bf16");
  return MMVQA_OK;
}

template <int BN, int STAGES, int KPS = 1>
static int launch_tc_major(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  if (a->a_trans && a->b_trans) return launch_tc<BN, true, true, STAGES, KPS>(a, ep, st);
  if (a->a_trans) return launch_tc<BN, true, false, STAGES, KPS>(a, ep, st);
  if (a->b_trans) return launch_tc<BN, false, true, STAGES, KPS>(a, ep, st);
  return launch_tc<BN, false, false, STAGES, KPS>(a, ep, st);
}


}  // namespace mmvqa
