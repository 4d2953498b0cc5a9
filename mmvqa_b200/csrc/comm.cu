// comm.cu -- in-switch all-reduce of a gradient bucket over NVLink 5 / NVSwitch (collective 1).
//
// The bucket lives in SYMMETRIC memory (same allocation on every rank, bound to one multicast object; the host side gets
// the multicast address and the per-rank signal pads from torch.distributed._symmetric_memory, which is plumbing).  The
// kernel is two-shot and in place:
//   1. every rank's CTA b synchronises with CTA b of every peer (signal pads, CAS 0 -> 1 / 1 -> 0, system scope);
//   2. rank r reduces slice r of the bucket: multimem.ld_reduce pulls the 16-byte vector from ALL ranks and the NVSwitch
//      adds them (fp32 accumulation for bf16x2), multimem.st broadcasts the sum back into every rank's copy;
//   3. the CTAs synchronise again: every slice has landed everywhere.
// Per rank the links carry 2 * bytes / world instead of the 2 * bytes * (world - 1) / world of a ring, and the kernel
// takes a handful of CTAs instead of NCCL's channels -- it shares the SMs with a latency-bound backward pass.
#include "common.cuh"

namespace mmvqa {

__device__ __forceinline__ void signal_put(uint32_t* addr) {
  uint32_t old;
  long long t0 = clock64();
  do {
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], 0, 1;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 0u && clock64() - t0 > 20000000000LL) __trap();      // a peer never consumed the previous signal
  } while (old != 0u);
}
__device__ __forceinline__ void signal_wait(uint32_t* addr) {
  uint32_t old;
  long long t0 = clock64();
  do {
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], 1, 0;" : "=r"(old) : "l"(addr) : "memory");
    if (old != 1u && clock64() - t0 > 20000000000LL) __trap();      // a peer never arrived
  } while (old != 1u);
}
// CTA b of this rank meets CTA b of every rank (one signal-pad word per (CTA, sender))
__device__ __forceinline__ void cta_barrier_all_ranks(uint32_t* const* pads, int rank, int world) {
  if ((int)threadIdx.x < world) {
    signal_put(pads[threadIdx.x] + (size_t)blockIdx.x * world + rank);
    signal_wait(pads[rank] + (size_t)blockIdx.x * world + threadIdx.x);
  }
}

template <bool BF16>
__global__ void __launch_bounds__(512) multimem_allreduce_kernel(char* __restrict__ mc, uint32_t* const* __restrict__ pads,
                                                                int rank, int world, long long n16) {
  cta_barrier_all_ranks(pads, rank, world);
  __syncthreads();
  const long long per = (n16 + world - 1) / world;
  const long long lo = per * rank, hi = (lo + per < n16) ? lo + per : n16;
  // a switch round trip is a few microseconds: UNROLL independent 16-byte reductions per thread are in flight before the
  // first result is needed (latency x bandwidth of the links, not SM count, sets the CTA budget)
  constexpr int UNROLL = 8;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += step * UNROLL) {
    uint32_t r[UNROLL][4];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * step;
      if (i < hi) {
        char* p = mc + i * 16;
        if (BF16)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "l"(p) : "memory");
        else
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(r[u][0]), "=r"(r[u][1]), "=r"(r[u][2]), "=r"(r[u][3]) : "l"(p) : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = i0 + u * step;
      if (i < hi)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc + i * 16), "r"(r[u][0]), "r"(r[u][1]),
                     "r"(r[u][2]), "r"(r[u][3]) : "memory");
    }
  }
  __threadfence_system();
  __syncthreads();
  cta_barrier_all_ranks(pads, rank, world);
}

}  // namespace mmvqa

using namespace mmvqa;

extern "C" {

int mmvqa_multimem_allreduce(void* multicast_ptr, const void* signal_pads_dev, int rank, int world, int64_t nbytes, int dtype,
                             int max_ctas, mmvqa_stream_t stream) {
  MMVQA_REQUIRE(multicast_ptr && signal_pads_dev && world >= 1 && world <= 32 && rank >= 0 && rank < world,
                "multimem_allreduce: bad communicator arguments");
  MMVQA_REQUIRE(nbytes >= 0 && nbytes % 16 == 0 && (reinterpret_cast<uintptr_t>(multicast_ptr) & 15) == 0,
                "multimem_allreduce: the buffer must be a multiple of 16 bytes and 16-byte aligned");
  MMVQA_REQUIRE(dtype == MMVQA_F32 || dtype == MMVQA_BF16, "multimem_allreduce: bad dtype %d", dtype);
  if (nbytes == 0) return MMVQA_OK;
  const long long n16 = nbytes / 16;
  int ctas = max_ctas > 0 ? max_ctas : 16;
  if (ctas > 64) ctas = 64;                              // signal pad: ctas * world words
  const long long want = (n16 / world + 511) / 512;
  if (want < ctas) ctas = want < 1 ? 1 : (int)want;
  char* mc = reinterpret_cast<char*>(multicast_ptr);
  uint32_t* const* pads = reinterpret_cast<uint32_t* const*>(signal_pads_dev);
  if (dtype == MMVQA_BF16)
    multimem_allreduce_kernel<true><<<ctas, 512, 0, as_stream(stream)>>>(mc, pads, rank, world, n16);
  else
    multimem_allreduce_kernel<false><<<ctas, 512, 0, as_stream(stream)>>>(mc, pads, rank, world, n16);
  MMVQA_LAUNCHED("multimem_allreduce");
  return MMVQA_OK;
}

}  // extern "C"
