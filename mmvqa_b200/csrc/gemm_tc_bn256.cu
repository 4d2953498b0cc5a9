#include "gemm_tc_kernel.cuh"

namespace mmvqa {

int launch_tc_bn256(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  if (kps == 2) return launch_tc_major<256, 2, 2>(a, ep, st);
  if (kps != 1) return set_err(MMVQA_ERR_ARG, "gemm(bf16): no %d-k-block stage for this tile", kps);
  switch (stages) {
    case 2: return launch_tc_major<256, 2>(a, ep, st);
    case 4: return launch_tc_major<256, 4>(a, ep, st);
    default: return set_err(MMVQA_ERR_ARG, "gemm(bf16): no %d-stage kernel for this tile", stages);
  }
}

}  // namespace mmvqa
