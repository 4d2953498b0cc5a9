// gemm.cu -- mmvqa_gemm: argument validation and dispatch to the fp32 SIMT kernel (dtype F32) or the
// tcgen05 / TMEM / TMA kernel (dtype BF16).  There is no fallback between the two: a bf16 problem
// that violates the TMA alignment rules is an error, not a silent SIMT run.
#include "gemm_common.cuh"
#include <stdlib.h>

using namespace mmvqa;

extern "C" int mmvqa_gemm(const mmvqa_gemm_args* a, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(a != nullptr, "gemm: null args");
  MMVQA_REQUIRE(a->dtype == MMVQA_F32 || a->dtype == MMVQA_BF16, "gemm: bad dtype %d", a->dtype);
  MMVQA_REQUIRE(a->c_dtype == MMVQA_F32 || a->c_dtype == MMVQA_BF16, "gemm: bad c_dtype %d", a->c_dtype);
  MMVQA_REQUIRE(a->M >= 0 && a->N >= 0 && a->K >= 0, "gemm: negative shape");
  MMVQA_REQUIRE(a->epilogue >= MMVQA_EPI_STORE && a->epilogue <= MMVQA_EPI_DACT_SCALE, "gemm: bad epilogue %d", a->epilogue);
  MMVQA_REQUIRE(a->act >= MMVQA_ACT_NONE && a->act <= MMVQA_ACT_RELU, "gemm: bad act %d", a->act);
  const int batch = a->batch < 1 ? 1 : a->batch;
  const int split_k = a->split_k < 1 ? 1 : a->split_k;
  if (a->M == 0 || a->N == 0) return MMVQA_OK;
  MMVQA_REQUIRE(a->A && a->B, "gemm: null operand");
  MMVQA_REQUIRE(a->lda >= (a->a_trans ? a->M : a->K), "gemm: lda %lld too small", (long long)a->lda);
  MMVQA_REQUIRE(a->ldb >= (a->b_trans ? a->N : a->K), "gemm: ldb %lld too small", (long long)a->ldb);
  if (a->epilogue != MMVQA_EPI_ACT_ROWSUM) {
    MMVQA_REQUIRE(a->C != nullptr && a->ldc >= a->N, "gemm: bad C / ldc");
  } else {
    MMVQA_REQUIRE(a->rowsum_out != nullptr, "gemm: EPI_ACT_ROWSUM needs rowsum_out");
  }
  if (a->epilogue == MMVQA_EPI_RESIDUAL || a->epilogue == MMVQA_EPI_DACT)
    MMVQA_REQUIRE(a->aux_in != nullptr && a->ld_aux_in >= a->N, "gemm: epilogue %d needs aux_in", a->epilogue);
  if (a->epilogue == MMVQA_EPI_DACT_SCALE) MMVQA_REQUIRE(a->rowscale != nullptr, "gemm: EPI_DACT_SCALE needs rowscale");
  if (a->rowscale != nullptr)
    MMVQA_REQUIRE(a->epilogue == MMVQA_EPI_DACT_SCALE || a->epilogue == MMVQA_EPI_STORE, "gemm: rowscale only with EPI_STORE / EPI_DACT_SCALE");
  if (a->aux_out) MMVQA_REQUIRE(a->ld_aux_out >= a->N, "gemm: bad ld_aux_out");
  if (a->colsum_out) MMVQA_REQUIRE(a->epilogue != MMVQA_EPI_ACT_ROWSUM && split_k == 1, "gemm: colsum_out needs a storing epilogue and split_k == 1");
  if (a->accumulate || split_k > 1) {
    const bool slabs = !a->accumulate && split_k > 1 && a->c_split_stride > 0;
    MMVQA_REQUIRE(a->c_dtype == MMVQA_F32 && a->epilogue == MMVQA_EPI_STORE && (a->accumulate || split_k == 1 || slabs),
                  "gemm: split_k / accumulate need an fp32 C, EPI_STORE and accumulate != 0 or c_split_stride > 0");
    if (slabs) MMVQA_REQUIRE(a->c_split_stride >= (int64_t)(a->M - 1) * a->ldc + a->N && batch == 1 && a->rowscale == nullptr,
                             "gemm: split-K slabs need batch == 1, no rowscale and c_split_stride >= one C matrix");
  }
  MMVQA_REQUIRE(a->dropout_p >= 0.0f && a->dropout_p < 1.0f, "gemm: dropout_p must be in [0,1)");
  if (batch > 1 && !a->accumulate && a->epilogue != MMVQA_EPI_ACT_ROWSUM)
    MMVQA_REQUIRE(a->c_batch_stride > 0, "gemm: batched store needs c_batch_stride");

  mmvqa_gemm_args v = *a;
  v.batch = batch;
  v.split_k = split_k;
  EpiParams ep;
  ep.M = a->M; ep.N = a->N; ep.K = a->K;
  ep.C = a->C; ep.ldc = a->ldc; ep.c_bf16 = (a->c_dtype == MMVQA_BF16);
  ep.bias = a->bias; ep.epilogue = a->epilogue; ep.act = a->act;
  ep.aux_in = a->aux_in; ep.ld_aux_in = a->ld_aux_in;
  ep.aux_out = a->aux_out; ep.ld_aux_out = a->ld_aux_out;
  ep.rowsum_out = a->rowsum_out; ep.colsum_out = a->colsum_out; ep.rowscale = a->rowscale; ep.scale = a->scale;
  ep.accumulate = a->accumulate; ep.split_k = split_k; ep.batch = batch;
  ep.c_batch_stride = a->c_batch_stride;
  ep.c_split_stride = (!a->accumulate && split_k > 1) ? a->c_split_stride : 0;
  ep.dropout_p = a->dropout_p; ep.dropout_seed = a->dropout_seed; ep.seed_ctr = g_seed_ctr;
  ep.trace = reinterpret_cast<unsigned long long*>(a->trace);
  static const int no_stage_env = getenv("MMVQA_TC_NO_STAGE") != nullptr ? atoi(getenv("MMVQA_TC_NO_STAGE")) : 0;
  ep.no_stage = no_stage_env;
  ep.serf_tab = 0;               // decided by the tcgen05 launcher
  if (a->dtype == MMVQA_F32) return gemm_simt_f32(&v, ep, as_stream(stream));
  int sm = mmvqa_device_sm();
  if (sm < 0) return sm;
  if (sm / 10 != 10) return set_err(MMVQA_ERR_ARCH, "gemm(bf16): tcgen05 path needs sm_100, device is sm_%d", sm);
  static const bool no_vistok = getenv("MMVQA_NO_VISTOK") != nullptr;   // A/B switch for tuning runs
  if (!no_vistok && vistok_applicable(&v)) return vistok_launch(&v, ep, as_stream(stream));
  return gemm_tc_bf16(&v, ep, as_stream(stream));
}
