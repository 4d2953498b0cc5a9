// gemm_simt.cu -- fp32 SIMT GEMM (FFMA, fp32 accumulate): the fp32-accurate validation path of
// mmvqa_gemm (dtype = MMVQA_F32).  tcgen05 has no true-fp32 MMA (tf32 only), so the path that must
// reproduce the reference's fp32 argmax runs on the CUDA cores.  64x64x16 tiles, 4x4 micro-tiles,
// arbitrary strides / transposes / batch / split-K; same epilogues as the tensor-core kernel.
#include "gemm_common.cuh"

namespace mmvqa {

constexpr int ST_BM = 64, ST_BN = 64, ST_BK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, int64_t sam, int64_t sak,
                                                        int64_t a_bstride, const float* __restrict__ B, int64_t sbn,
                                                        int64_t sbk, int64_t b_bstride, EpiParams p, int a_kfast,
                                                        int b_kfast) {
  __shared__ float As[ST_BK][ST_BM + 4];
  __shared__ float Bs[ST_BK][ST_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4 (m) x 4 (n)
  const int m0 = blockIdx.y * ST_BM, n0 = blockIdx.x * ST_BN;
  const int bz = blockIdx.z / p.split_k, ks = blockIdx.z % p.split_k;
  const int kblocks = (p.K + ST_BK - 1) / ST_BK;
  const int kb_per = (kblocks + p.split_k - 1) / p.split_k;
  const int kb0 = ks * kb_per, kb1 = min(kblocks, kb0 + kb_per);
  A += (int64_t)bz * a_bstride;
  B += (int64_t)bz * b_bstride;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int kb = kb0; kb < kb1; ++kb) {
    const int k0 = kb * ST_BK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int idx = tid + i * 256;
      int mm, kk;
      if (a_kfast) { kk = idx & 15; mm = idx >> 4; } else { mm = idx & 63; kk = idx >> 6; }
      int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < p.M && gk < p.K) ? __ldg(A + gm * sam + gk * sak) : 0.0f;
      int nn;
      if (b_kfast) { kk = idx & 15; nn = idx >> 4; } else { nn = idx & 63; kk = idx >> 6; }
      int gn = n0 + nn;
      gk = k0 + kk;
      Bs[kk][nn] = (gn < p.N && gk < p.K) ? __ldg(B + gn * sbn + gk * sbk) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < ST_BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (kb0 >= kb1 && ks != 0 && p.c_split_stride == 0) return;  // empty split (a slab still gets its zeros)
  const int64_t c_split_off = (int64_t)ks * p.c_split_stride;
  const bool first = (ks == 0);
  float cs[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    float rs = 0.0f;
    if (m < p.M) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        if (n < p.N) {
          const float o = epi_element<float>(p, bz, m, n, acc[i][j], first, c_split_off);
          rs += o;
          cs[j] += o;
        }
      }
    }
    if (p.epilogue == MMVQA_EPI_ACT_ROWSUM) {
      // the 16 tx-threads of one row are consecutive lanes of a half warp
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, o);
      if (tx == 0 && m < p.M) atomicAdd(p.rowsum_out + (int64_t)bz * p.M + m, rs * p.scale);
    }
  }
  if (p.colsum_out && p.epilogue != MMVQA_EPI_ACT_ROWSUM) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.N) atomicAdd(p.colsum_out + n, cs[j]);
    }
  }
}

int gemm_simt_f32(const mmvqa_gemm_args* a, const EpiParams& ep, cudaStream_t st) {
  MMVQA_REQUIRE(a->c_dtype == MMVQA_F32, "gemm(f32): c_dtype must be f32");
  int64_t sam = a->a_trans ? 1 : a->lda, sak = a->a_trans ? a->lda : 1;
  int64_t sbn = a->b_trans ? 1 : a->ldb, sbk = a->b_trans ? a->ldb : 1;
  dim3 grid((a->N + ST_BN - 1) / ST_BN, (a->M + ST_BM - 1) / ST_BM, a->batch * a->split_k);
  MMVQA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "gemm(f32): grid too large (M=%d batch=%d)", a->M, a->batch);
  gemm_simt_kernel<<<grid, 256, 0, st>>>((const float*)a->A, sam, sak, a->a_batch_rows * a->lda, (const float*)a->B, sbn,
                                         sbk, a->b_batch_rows * a->ldb, ep, !a->a_trans, !a->b_trans);
  MMVQA_LAUNCHED("gemm_simt_f32");
  return MMVQA_OK;
}

}  // namespace mmvqa
