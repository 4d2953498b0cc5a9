// similarity.cu -- caption-similarity masks for the SupCon pre-training step on the GPU.
// models/SupConLoss/supcon_utils.py:110-138 builds the [bsz, bsz] Jaccard mask with two nested Python loops over
// word sets (O(bsz^2) set intersections on the host: the step bottleneck at bsz = 1024).
// Here every document is a sorted array of unique word ids (built once on the host, O(words)); one thread per
// (caption, augmentation) pair merges the two arrays and counts the intersection.  Integer work + one IEEE
// division: bit-exact with the reference's float(len(intersection)) / len(union) stored into a float32 tensor.
#include "common.cuh"

namespace mmvqa {

constexpr int JAC_ROWS = 4;      // caption rows per CTA (their id arrays are staged in shared memory)

__global__ void __launch_bounds__(256) jaccard_mask_kernel(const int* __restrict__ ids_a, const int* __restrict__ len_a,
                                                           const int* __restrict__ ids_b, const int* __restrict__ len_b,
                                                           float* __restrict__ mask, int na, int nb, int lmax) {
  extern __shared__ int sm_ids[];          // [JAC_ROWS][lmax]
  const int r0 = blockIdx.x * JAC_ROWS;
  pdl_wait();
  pdl_trigger();
  for (int i = threadIdx.x; i < JAC_ROWS * lmax; i += blockDim.x) {
    const int r = r0 + i / lmax;
    sm_ids[i] = r < na ? ids_a[(int64_t)r * lmax + (i % lmax)] : 0;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < JAC_ROWS * nb; idx += blockDim.x) {
    const int rr = idx / nb, c2 = idx - rr * nb;
    const int c1 = r0 + rr;
    if (c1 >= na) break;
    float out;
    if (c1 == c2) {
      out = 1.0f;                                            // supcon_utils.py:116-117
    } else {
      const int* a = sm_ids + rr * lmax;
      const int* b = ids_b + (int64_t)c2 * lmax;
      const int la = len_a[c1], lb = len_b[c2];
      int i = 0, j = 0, inter = 0;
      while (i < la && j < lb) {
        const int x = a[i], y = __ldg(b + j);
        inter += (x == y);
        i += (x <= y);
        j += (y <= x);
      }
      const int uni = la + lb - inter;
      out = uni != 0 ? (float)((double)inter / (double)uni) : 0.0f;   // :133-138
    }
    mask[(int64_t)c1 * nb + c2] = out;
  }
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" int mmvqa_jaccard_mask(const int* ids_a, const int* len_a, const int* ids_b, const int* len_b, float* mask,
                                  int na, int nb, int lmax, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(ids_a && len_a && ids_b && len_b && mask, "jaccard_mask: null pointer");
  MMVQA_REQUIRE(na >= 0 && nb >= 0 && lmax >= 1, "jaccard_mask: bad shape");
  if (na == 0 || nb == 0) return MMVQA_OK;
  const size_t smem = sizeof(int) * JAC_ROWS * (size_t)lmax;
  MMVQA_REQUIRE(smem <= 48 * 1024, "jaccard_mask: documents of more than %d unique words are not supported", 48 * 1024 / 4 / JAC_ROWS);
  MMVQA_CUDA(launch_pdl(jaccard_mask_kernel, dim3((na + JAC_ROWS - 1) / JAC_ROWS), dim3(256), smem, as_stream(stream), ids_a,
                        len_a, ids_b, len_b, mask, na, nb, lmax));
  MMVQA_LAUNCHED("jaccard_mask");
  return MMVQA_OK;
}
