// rf_cluster.cuh -- PTX helpers shared by the cluster kernels of the RealFormer encoder (rf_encoder.cu: the whole forward
// in one launch; rf_attn_block.cu: the attention block of one layer, backward): 4-D TMA loads, cluster ranks / barriers,
// remote mbarrier arrivals, distributed-shared-memory stores.
#pragma once
#include "gemm_tc_kernel.cuh"
#include "attention_tc.cuh"

namespace mmvqa {

constexpr int RFC_HEADS = 8;              // CTAs per cluster = attention heads (models/mmbert.py:100)

// ---------------------------------------------------------------------------------------------------------------
// PTX helpers that gemm_tc_kernel.cuh does not have
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
// one arrival on the same barrier of every CTA of the cluster (release at cluster scope: what this CTA wrote before --
// global memory and peers' shared memory -- is visible to whoever observes the completed phase with acquire)
__device__ __forceinline__ void cluster_arrive_all(uint32_t local_bar) {
#pragma unroll
  for (uint32_t r = 0; r < RFC_HEADS; ++r) {
    const uint32_t remote = map_to_cta(local_bar, r);
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  }
}
__device__ __forceinline__ bool mbar_try_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_cluster(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t remote, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(remote), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
// after a bar.sync of the 256 compute threads: 8 lanes publish "this CTA reached the sync point" to the 8 CTAs of the
// cluster.  One release fence per arriving lane (cumulative over what the other threads wrote before the bar.sync),
// then a relaxed remote arrival.
__device__ __forceinline__ void cluster_publish(uint32_t local_bar, int ctid) {
  if (ctid < RFC_HEADS) {
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const uint32_t remote = map_to_cta(local_bar, (uint32_t)ctid);
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
  }
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}


}  // namespace mmvqa
