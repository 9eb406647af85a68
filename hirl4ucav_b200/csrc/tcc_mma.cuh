// tcc_mma.cuh -- tcgen05 building blocks of the tensor-core cluster kernel (tcc_update.cuh): shared-memory operand
// tiles in the canonical SWIZZLE_128B layout, descriptors for K-major AND MN-major use of the same bytes, 3xTF32
// product issue, bulk copies, cluster / DSMEM primitives.  Everything here is exercised stand-alone by
// tools/tcc_probe.cu against fp64 on the host.
//
// Operand tile ("block"): R rows x 32 fp32 (one 128-byte swizzle row per tile row), rows grouped in 8-row atoms of
// 1024 bytes, the 16-byte chunk index XOR-ed with (row & 7).  A matrix [R][K] is a sequence of K/32 blocks.
//   * read K-major (reduction index = the 32-float direction): descriptor SBO = 1024 (next 8-row atom); an MMA k-step
//     (8 tf32) advances the start address by 32 bytes inside the swizzle row.
//   * read MN-major (reduction index = the ROW direction, M/N index = the 32-float direction): the very same bytes are
//     the canonical MN-major SWIZZLE_128B layout ((8 chunks, n atoms), (8 k, k atoms)) with LBO = block stride
//     (R * 128 bytes: next 32 M/N indices) and SBO = 1024 (next 8 reduction indices); a k-step advances by 1024 bytes.
//   That identity is what lets the forward operand H1 [rows][256] serve unchanged as the operand of the weight-gradient
//   contraction over the rows (dW2 = dZ2^T . H1) -- no transposed copy.
#pragma once
#include "tc_gemm.cuh"

namespace hirl {
namespace tcc {

using tc::smem_u32;

__device__ __forceinline__ uint32_t sw128_off(int row, int kk) {   // byte offset of element (row, kk) inside a block, kk in [0, 32)
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((kk >> 2) ^ (row & 7)) << 4) | ((kk & 3) << 2)));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// ---- descriptors ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {       // K-major, SWIZZLE_128B, 8-row atoms 1024 B apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo_bytes) {   // MN-major, SWIZZLE_128B: LBO = next 32 M/N indices, SBO = next 8 k
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::tf32, fp32 accumulate; a_mn / b_mn: operand is MN-major
__device__ __forceinline__ uint32_t idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t id, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(id), "r"(accumulate)
        : "memory");
}

// One fp32-accurate contraction step set: D_main (+)= Ahi.Bhi ; D_corr (+)= Alo.Bhi + Ahi.Blo over `ksteps` k-steps.
// a_step / b_step: descriptor advance per k-step in 16-byte units (2 for K-major, 64 for MN-major).
// first: the accumulators are overwritten by the first k-step.  Issued by ONE thread.
__device__ __forceinline__ void mma3(uint32_t d_main, uint32_t d_corr, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, int ksteps,
                                     uint32_t a_step, uint32_t b_step, uint32_t id, bool first) {
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t ao = (uint64_t)(a_step * k), bo = (uint64_t)(b_step * k);
        const uint32_t acc = (first && k == 0) ? 0u : 1u;
        mma_ss(d_main, a_hi + ao, b_hi + bo, id, acc);
        mma_ss(d_corr, a_lo + ao, b_hi + bo, id, acc);
        mma_ss(d_corr, a_hi + ao, b_lo + bo, id, 1u);
    }
}

// ---- bulk copy global -> shared with mbarrier completion (cp.async.bulk, SASS UBLKCP) ----------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same copy delivered to the same CTA-relative offsets (data and mbarrier) of every CTA in cta_mask (SASS UBLKCP ... .MULTICAST)
__device__ __forceinline__ void bulk_g2s_mc(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}

// ---- cluster / DSMEM -----------------------------------------------------------------------------------------------
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
__device__ __forceinline__ void cluster_sync() {   // all threads of all CTAs; release/acquire at cluster scope
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// split-phase form: arrive (release) right after producing, wait (acquire) right before consuming
__device__ __forceinline__ void cluster_arrive() {
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
    __syncwarp();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// remote 16-byte store that signals the DESTINATION CTA's mbarrier with 16 transaction bytes (no cluster barrier needed)
__device__ __forceinline__ void st_async_f32x4(uint32_t remote_addr, float a, float b, float c, float d, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr),
                 "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d)), "r"(remote_bar)
                 : "memory");
}
// generic-proxy global writes (made visible by a barrier) -> later async-proxy reads (bulk copies)
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void st_dsmem_f32x2(uint32_t addr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void st_dsmem_f32x4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace tcc
}  // namespace hirl
