// gemm_common.cuh -- epilogue parameters shared by the SIMT fp32 GEMM and the tcgen05 bf16 GEMM.
#pragma once
#include "common.cuh"

namespace mmvqa {

struct EpiParams {
  int M, N, K;
  void* C; int64_t ldc; int c_bf16;          // output
  const float* bias;
  int epilogue, act;
  const void* aux_in; int64_t ld_aux_in;      // storage type == operand type
  void* aux_out; int64_t ld_aux_out;
  float* rowsum_out; float* colsum_out; const float* rowscale; float scale;
  int accumulate, split_k, batch;
  int64_t c_batch_stride;
  int64_t c_split_stride;
  float dropout_p; unsigned long long dropout_seed; const unsigned long long* seed_ctr;
  unsigned long long* trace;
  int no_stage;                                // 1: direct per-row stores in the tcgen05 epilogue (A/B tuning: MMVQA_TC_NO_STAGE)
  int serf_tab;                                // 1: the launch carries TC_SERF_TAB_BYTES of shared memory for the SERF table
};

// one output element; AUX = storage type of aux_in / aux_out.  `first_split` gates bias so that
// split-K partial sums add it once.  Returns the stored value (the activation for EPI_ACT_ROWSUM).
template <typename AUX>
__device__ __forceinline__ float epi_element(const EpiParams& p, int bz, int m, int n, float acc, bool first_split,
                                             int64_t c_split_off = 0) {
  float v = acc + ((p.bias && first_split) ? __ldg(p.bias + n) : 0.0f);
  float out;
  switch (p.epilogue) {
    case MMVQA_EPI_ACT:
      if (p.aux_out) reinterpret_cast<AUX*>(p.aux_out)[(int64_t)m * p.ld_aux_out + n] = from_f<AUX>(v);
      out = act_rt(p.act, v);
      break;
    case MMVQA_EPI_RESIDUAL: {
      if (p.dropout_p > 0.0f) {
        uint32_t thr = (uint32_t)(p.dropout_p * 4294967296.0);
        v = hash32(seed_eff(p.dropout_seed, p.seed_ctr), (uint64_t)m * (uint64_t)p.N + (uint64_t)n) >= thr ? v / (1.0f - p.dropout_p) : 0.0f;
      }
      out = v + to_f(reinterpret_cast<const AUX*>(p.aux_in)[(int64_t)m * p.ld_aux_in + n]);
      break;
    }
    case MMVQA_EPI_DACT:
      out = v * dact_rt(p.act, to_f(reinterpret_cast<const AUX*>(p.aux_in)[(int64_t)m * p.ld_aux_in + n]));
      break;
    case MMVQA_EPI_ACT_ROWSUM:
      if (p.aux_out)   // act'(.) kept for the backward pass: [batch, M, ld_aux_out]
        reinterpret_cast<AUX*>(p.aux_out)[((int64_t)bz * p.M + m) * p.ld_aux_out + n] = from_f<AUX>(dact_rt(p.act, v));
      return act_rt(p.act, v);
    case MMVQA_EPI_DACT_SCALE:
      out = dact_rt(p.act, v) * __ldg(p.rowscale + (int64_t)bz * p.M + m) * p.scale;
      break;
    default:
      out = p.rowscale ? v * __ldg(p.rowscale + (int64_t)bz * p.M + m) * p.scale : v;
  }
  int64_t off = (int64_t)bz * p.c_batch_stride + c_split_off + (int64_t)m * p.ldc + n;
  if (p.accumulate)
    atomicAdd(reinterpret_cast<float*>(p.C) + off, out);
  else if (p.c_bf16)
    reinterpret_cast<__nv_bfloat16*>(p.C)[off] = __float2bfloat16_rn(out);
  else
    reinterpret_cast<float*>(p.C)[off] = out;
  return out;
}

int gemm_simt_f32(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
int gemm_tc_bf16(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);
bool vistok_applicable(const mmvqa_gemm_args* a);
int vistok_launch(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st);

}  // namespace mmvqa
