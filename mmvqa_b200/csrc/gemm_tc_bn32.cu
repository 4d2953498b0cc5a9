#include "gemm_tc_kernel.cuh"

namespace mmvqa {

int launch_tc_bn32(int stages, int kps, const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  if (kps == 2) return a->a_trans ? launch_tc<32, true, false, 2, 2>(a, ep, st) : launch_tc<32, false, false, 2, 2>(a, ep, st);
  if (kps == 4) return a->a_trans ? launch_tc<32, true, false, 2, 4>(a, ep, st) : launch_tc<32, false, false, 2, 4>(a, ep, st);
  switch (stages) {
    case 2: return a->a_trans ? launch_tc<32, true, false, 2>(a, ep, st) : launch_tc<32, false, false, 2>(a, ep, st);
    case 4: return a->a_trans ? launch_tc<32, true, false, 4>(a, ep, st) : launch_tc<32, false, false, 4>(a, ep, st);
    case 8: return a->a_trans ? launch_tc<32, true, false, 8>(a, ep, st) : launch_tc<32, false, false, 8>(a, ep, st);
    default: return set_err(MMVQA_ERR_ARG, "gemm(bf16): no %d-stage kernel for this tile", stages);
  }
}

}  // namespace mmvqa
