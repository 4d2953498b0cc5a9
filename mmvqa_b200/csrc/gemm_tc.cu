// gemm_tc.cu -- bf16 GEMM on the 5th-generation tensor cores of sm_100a:
//   TMA (cp.async.bulk.tensor, 128-byte swizzle) -> shared-memory ring -> tcgen05.mma (one issuing
//   thread, cta_group::1, 128 x BN x 16 per instruction) -> fp32 accumulator in TMEM -> tcgen05.ld ->
//   fused epilogue (bias / SERF / GELU / ReLU / residual(+dropout) / act' / pooled row-sum) -> global.
// Warp roles (192 threads): warp 0 = TMEM allocator + TMA producer, warp 1 = barrier init + MMA
// issuer, warps 2..5 = epilogue (each owns the 32 TMEM lanes  (warp_idx & 3) * 32 ...).
// Both operands may be K-major (contraction dim contiguous: activations x nn.Linear weights) or
// MN-major (the transposed reads that dgrad / wgrad need) -- the latter use the MN-major UMMA
// shared-memory descriptors, so no transposed copies of weights or activations are ever made.
#include "gemm_common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace mmvqa {

constexpr int TC_BM = 128, TC_BK = 64;

int launch_tc_bn32(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
int launch_tc_bn64(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
int launch_tc_bn128(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
int launch_tc_bn256(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
static bool chunkable(const mmvqa_gemm_args* a) {
  const bool a_ok = a->a_trans ? (a->M % 64 == 0) : (a->K % 64 == 0);
  const bool b_ok = a->b_trans ? (a->N % 64 == 0) : (a->K % 64 == 0);
  return a_ok && b_ok;
}

// ---------------------------------------------------------------------------------
// host: tensor maps + dispatch
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 3-D map over a stored [rows, inner] bf16 matrix with `nbatch` slabs `batch_rows` rows apart.
int tc_make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int nbatch,
                    int64_t batch_rows, int box_inner, int box_rows, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(MMVQA_ERR_CUDA, "gemm(bf16): cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0)
    return set_err(MMVQA_ERR_ARG, "gemm(bf16): operand %s needs a 16-byte aligned base and a leading dimension that is a multiple of 8 (ld=%lld)", what, (long long)ld);
  const bool batched = nbatch > 1 && batch_rows > 0;
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)(batched ? nbatch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batched ? batch_rows * ld * 2 : rows * ld * 2)};
  if (strides[1] % 16 != 0) strides[1] = (strides[1] + 15) / 16 * 16;
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(MMVQA_ERR_CUDA, "gemm(bf16): cuTensorMapEncodeTiled(%s) failed with %d (inner=%lld rows=%lld ld=%lld)", what, (int)r, (long long)inner, (long long)rows, (long long)ld);
  return MMVQA_OK;
}

int tc_make_map_chunked(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t ld, int nbatch,
                        int64_t batch_rows, int box_rows, int box_chunks, const char* what) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(MMVQA_ERR_CUDA, "gemm(bf16): cuTensorMapEncodeTiled unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 2) % 16 != 0 || inner % 64 != 0 || box_rows > 256 || box_chunks > 256 ||
      box_chunks < 1)
    return set_err(MMVQA_ERR_ARG, "gemm(bf16): operand %s cannot use the chunked map (ld=%lld inner=%lld)", what, (long long)ld,
                   (long long)inner);
  const bool batched = nbatch > 1 && batch_rows > 0;
  cuuint64_t dims[4] = {64, (cuuint64_t)rows, (cuuint64_t)(inner / 64), (cuuint64_t)(batched ? nbatch : 1)};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, 128, (cuuint64_t)(batched ? batch_rows * ld * 2 : rows * ld * 2)};
  if (strides[2] % 16 != 0) strides[2] = (strides[2] + 15) / 16 * 16;
  cuuint32_t box[4] = {64, (cuuint32_t)box_rows, (cuuint32_t)box_chunks, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(MMVQA_ERR_CUDA, "gemm(bf16): cuTensorMapEncodeTiled(%s, chunked) failed with %d (inner=%lld rows=%lld ld=%lld box %d x %d)", what, (int)r, (long long)inner, (long long)rows, (long long)ld, box_rows, box_chunks);
  return MMVQA_OK;
}

static int env_int(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}

int gemm_tc_bf16(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  // Tile width: the widest BN that still gives the chip a full wave of CTAs (two for BN = 256).  Epilogues that
  // evaluate an activation per element are bound by the epilogue warps, not the MMA: they take BN <= 128.
  // MN-major B needs BN % 64 == 0.
  // Ring depth: a problem of one wave or less takes the deepest ring that fits (up to 8 stages, ~190 KB);
  // multi-wave problems take a 2-stage ring so that 2-3 CTAs share an SM and the epilogue of one CTA overlaps the
  // main loop of the others (measured at M = 28672: FF1 + SERF 466 -> 270 us, plain GEMMs unchanged or faster).  MMVQA_TC_BN / MMVQA_TC_STAGES override the choice (tuning only).
  const int sms = num_sms();
  const int64_t mt = (a->M + TC_BM - 1) / TC_BM;
  const int64_t z = (int64_t)a->batch * a->split_k;
  const bool heavy = (a->act != MMVQA_ACT_NONE && a->act != MMVQA_ACT_RELU) &&
                     (a->epilogue == MMVQA_EPI_ACT || a->epilogue == MMVQA_EPI_ACT_ROWSUM ||
                      a->epilogue == MMVQA_EPI_DACT_SCALE || a->epilogue == MMVQA_EPI_DACT);
  const int cand[4] = {256, 128, 64, 32};
  const int ncand = a->b_trans ? 3 : 4;
  int bn = cand[ncand - 1];
  for (int i = heavy ? 1 : 0; i < ncand; ++i) {
    const int b = cand[i];
    if (b > 32 && b / 2 >= a->N) continue;  // tile wider than twice the problem
    const int64_t ctas = mt * ((a->N + b - 1) / b) * z;
    if (ctas * 10 >= (int64_t)sms * 9 * (b == 256 ? 2 : 1)) { bn = b; break; }   // >= 90 % of a wave
  }
  static const int env_bn = env_int("MMVQA_TC_BN"), env_st = env_int("MMVQA_TC_STAGES");
  if (env_bn == 32 || env_bn == 64 || env_bn == 128 || env_bn == 256) {
    bn = env_bn;
    if (a->b_trans && bn < 64) bn = 64;
  }
  const int kblocks = ((a->K + TC_BK - 1) / TC_BK + a->split_k - 1) / a->split_k;
  const int64_t ctas = mt * ((a->N + bn - 1) / bn) * z;
  const int deep = bn == 256 ? 4 : (bn == 128 ? 6 : 8);
  int stages;
  if (kblocks <= 2) stages = 2;
  else if (ctas <= sms && kblocks > 4) stages = deep;
  else if (ctas > 2 * (int64_t)sms) stages = 2;   // multi-wave: 2-3 CTAs per SM, epilogue of one hides under the main loop of the others
  else stages = 4;
  if (env_st == 2 || env_st == 4 || env_st == 6 || env_st == 8) stages = env_st > deep ? deep : env_st;
  if (stages == 6 && bn != 128) stages = 4;
  if (stages == 8 && bn > 64) stages = deep;
  // Programmatic dependent launch only overlaps kernels whose CTAs fit on an SM together (and the side-branch
  // weight-gradient GEMMs share SMs with the main chain the same way): rings of <= ~96 KB, two CTAs per SM.
  if (pdl_enabled() && env_st == 0) {
    const int cap = bn == 128 ? 3 : (bn == 256 ? 2 : 4);
    if (stages > cap) stages = cap;
  }
  // k-blocks per ring stage.  A stage costs a CTA ~0.3-0.4 us whatever it carries (profiles/r02_tma_dsmem_ubench.txt,
  // profiles/r01_gemm_phase_trace.txt), so short-K problems -- the whole fine-tune step: K = 448 ... 3072 / split --
  // take two fat stages: 4 k-blocks per stage when one CTA per SM covers the grid, 2 when two CTAs must share an SM
  // (<= 96 KB rings).  Long-K multi-wave problems keep the thin ring with 2-3 CTAs per SM (epilogue under main loop).
  int kps = 1;
  static const int env_kps = env_int("MMVQA_TC_KPS");
  const int per_kb = TC_BM * TC_BK * 2 + bn * TC_BK * 2;
  const bool fits4 = 2 * 4 * per_kb <= 200 * 1024, fits2 = 2 * 2 * per_kb <= 200 * 1024;
  const bool fits2_pair = 2 * 2 * per_kb <= 100 * 1024;                  // two such CTAs on one SM
  // (wide tiles with few k-blocks -- the K = 448 weight-gradient GEMMs of the side branch -- keep the thin 96 KB ring: a
  // 128 KB ring takes the whole SM and, inside the step, delays the main-chain kernels that share it: 26.7 -> 36.2 us)
  if (chunkable(a) && kblocks >= 2 && kblocks <= 24 && (bn < 128 || kblocks >= 12)) {
    if (ctas <= sms) kps = (fits4 && kblocks >= 4) ? 4 : (fits2 ? 2 : 1);
    else if (ctas <= 2 * (int64_t)sms) kps = fits2_pair ? 2 : 1;         // one wave only with two CTAs per SM
    else kps = fits2 ? 2 : 1;
  }
  if (env_kps == 1) kps = 1;
  if ((env_kps == 2 || env_kps == 4) && chunkable(a) && kblocks >= 2) kps = (env_kps == 4 && fits4) ? 4 : (fits2 ? 2 : 1);
  if (kps > 1) stages = 2;
  switch (bn) {
    case 256: return launch_tc_bn256(stages, kps, a, ep, st);
    case 128: return launch_tc_bn128(stages, kps, a, ep, st);
    case 64: return launch_tc_bn64(stages, kps, a, ep, st);
    default: return launch_tc_bn32(stages, kps, a, ep, st);
  }
}

}  // namespace mmvqa
