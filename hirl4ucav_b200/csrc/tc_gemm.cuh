// tc_gemm.cuh -- fp32-accurate GEMM tiles on the 5th-gen tensor cores (tcgen05, sm_100a).
//
//   D[128 x 64] = A[128 x K] . B[64 x K]^T      A, B fp32, both K-major (reduction index contiguous)
//
// fp32 accuracy from TF32 tensor cores by the 3-product split (3xTF32): x = hi + lo with
// hi = x truncated to TF32 (what the tensor core reads from an fp32 word anyway) and lo = x - hi
// (exact in fp32); D += Ahi.Bhi + Alo.Bhi + Ahi.Blo, accumulated in fp32 in TMEM.  The dropped
// lo.lo term and the truncation of lo are ~2^-21 relative per product.
//
// One CTA of 256 threads per tile:
//   * all threads stage 32-wide K blocks: coalesced fp32 loads (prefetched one block ahead in
//     registers), split, store into the canonical K-major SWIZZLE_128B shared-memory layout
//     (8-row x 128-byte atoms, 16-byte chunks XOR-swizzled by the row index);
//   * one thread issues 12 tcgen05.mma.kind::tf32 per block (4 k-steps x 3 products) and commits
//     to the block's mbarrier, which gates re-use of that shared-memory stage (4 stages);
//   * the accumulator lives in 64 TMEM columns allocated once per kernel; the epilogue reads it
//     back with tcgen05.ld (each warp its own 32-lane quarter), adds the bias and stores fp32.
#pragma once
#ifndef HIRL_TC_DEBUG
#define HIRL_TC_DEBUG 0   // test-only knobs: 1 = no MMAs (staging only), 2 = no staging stores (MMAs only), 3 = no global loads
#endif
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>

namespace hirl {
namespace tc {
#if HIRL_TC_DEBUG == 4
__device__ long long g_tc_dbg[64];
#define TC_T(i) if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)&g_tc_dbg[i], (unsigned long long)(clock64() - t_));
#define TC_T0 long long t_ = clock64();
#define TC_TR t_ = clock64();
#else
#define TC_T(i)
#define TC_T0
#define TC_TR
#endif

constexpr int BM = 128, BN = 64, BK = 32;            // tile and K-block (32 fp32 = one 128-byte swizzle row)
constexpr int NSTAGE = 4;
constexpr int B_BYTES = BN * BK * 4;                  // 8 KB per part (hi, lo) per stage
constexpr int SMEM_BYTES = NSTAGE * 2 * B_BYTES + 1024;  // B ring + alignment slack; barriers live in static smem
constexpr int TMEM_COLS = 512;                        // 3 x 64 accumulator columns + 4 stages x 64 columns of A (hi|lo)
constexpr int THREADS_PER_TILE = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMEM allocation (one warp, whole CTA lifetime) -----------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a pipeline bug must fail loudly (trap -> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int tag = 0) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    bool reported = false;
    while (!mbar_try(bar, parity)) {
        const long long dt = clock64() - t0;
        if (dt > 4000000000LL && !reported) {   // ~2 s: report, keep waiting a little so every stuck warp reports, then trap
            reported = true;
            if ((threadIdx.x & 31) == 0)
                printf("hirl tc_gemm: mbarrier wait timed out (block %d warp %d tag %d parity %u)\n", blockIdx.x, threadIdx.x >> 5, tag, parity);
        }
        if (dt > 6000000000LL) __trap();
    }
}

// ---- descriptors -----------------------------------------------------------------------------
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFF) >> 4);          // start address  [0,14)
    // leading byte offset [16,30): unused for swizzled K-major operands, left 0
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset: next 8-row atom [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                           // layout: SWIZZLE_128B
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major, M=128, N=64
__device__ __forceinline__ uint32_t instr_desc() {
    return (1u << 4)                 // D format F32
           | (2u << 7)               // A format TF32
           | (2u << 10)              // B format TF32
           | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of (row, 16-byte chunk c) inside a [rows x 128 B] SWIZZLE_128B block
__device__ __forceinline__ uint32_t swz(int row, int c) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4)); }

__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// hi/lo split of four fp32 and store to the two shared-memory parts (shared-space addresses)
__device__ __forceinline__ void split_store(uint32_t hi_saddr, uint32_t lo_saddr, uint32_t off, float4 x) {
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
    sts128(hi_saddr + off, h);
    sts128(lo_saddr + off, l);
}

// One 128x64 output tile.  A: rows m0.. (valid while < M) of [M][K] (lda), B: rows n0.. of [N][K] (ldb);
// K-range [k_begin, k_end) in elements (multiples of 4; the tail of a 32-wide block is zero-filled).
// C[m][n] (ldc) = sum + bias[n].  `smem` must be 1024-byte aligned with tc::SMEM_BYTES available.
// All 256 threads must call.
//
// Operand placement: the tensor core reads B (64 rows) from shared memory and A (128 rows) from
// TENSOR MEMORY.  With both operands in shared memory a 128x64x8 tf32 MMA moves 6 KB through the
// 128 B/clk shared-memory port (48 cycles for 32 cycles of math, x3 products, plus the staging
// stores): measured 11.5 us per tile.  A-in-TMEM makes thread r the owner of row r: it loads its
// row's 16 fp32 of the K block from global, splits them and writes hi/lo to its TMEM lane with
// tcgen05.st; only B goes through shared memory (canonical K-major SWIZZLE_128B layout).
//
// Pipeline: 4-stage ring (B in smem, A in TMEM columns) guarded by full/empty mbarriers; warp 0's
// elected lane issues the MMAs of block kb-1 while everybody stages block kb, global loads are
// prefetched one block ahead in registers.
//
// Accuracy: the tensor core accumulates with truncation, so the error of a long accumulation chain
// grows with its length.  Two measures keep the result at fp32-FMA quality: (1) the large product
// Ahi.Bhi and the two small corrections (Alo.Bhi + Ahi.Blo) go to SEPARATE TMEM accumulators that
// are added in fp32 (round-to-nearest) in the epilogue, so the main chain has K/8 steps, not 3K/8;
// (2) K is processed in chunks of at most KCHUNK elements whose partial results are summed in
// registers with ordinary fp32 adds.
constexpr int KCHUNK = 512;
// TMEM columns: three accumulators (Ahi.Bhi | Alo.Bhi | Ahi.Blo), then the A ring (hi 32 | lo 32 per stage)
constexpr int COL_MAIN = 0, COL_CORR = BN, COL_CORR2 = 2 * BN, COL_A = 3 * BN;
constexpr int ISSUERS = 3;          // warps 0,1,2 each issue one of the three products (a single thread issuing all
                                    // twelve MMAs of a block costs ~70 cycles per MMA and was the critical path)
// Initialised ONCE per kernel (barriers_init); phases are tracked across tiles with the two running
// counters, the mbarrier objects are never re-initialised while the kernel runs.
struct Barriers {
    uint64_t full[NSTAGE], empty[NSTAGE], chunk;
    uint32_t blocks_done;   // K blocks staged so far by this CTA (all tiles)
    uint32_t chunks_done;   // chunk completions so far
};

// once per kernel, by one thread, followed by __syncthreads()
__device__ __forceinline__ void barriers_init(Barriers *bars) {
#pragma unroll
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(bars->full + s, THREADS_PER_TILE / 32); mbar_init(bars->empty + s, ISSUERS); }
    mbar_init(&bars->chunk, ISSUERS);
    bars->blocks_done = 0u;
    bars->chunks_done = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem desc]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __noinline__ void gemm_tile(const float *A, int lda, int M, const float *Bm, int ldb, int N, int K,
                                       int m0, int n0, int k_begin, int k_end, float *C, int ldc,
                                       const float *bias, uint32_t tmem_base, char *smem, Barriers *bars) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkb = (k_end - k_begin + BK - 1) / BK;
    constexpr int KB_PER_CHUNK = KCHUNK / BK;
    const int nchunks = (nkb + KB_PER_CHUNK - 1) / KB_PER_CHUNK;
    const int q = warp & 3, half = warp >> 2;
    const int arow = m0 + q * 32 + lane;                  // the A row (= TMEM lane) this thread owns
    const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int bc = tid & 7, br = tid >> 3;                // B staging: 16-byte chunk and row (0..31, +32)

    __syncthreads();                         // previous users of smem / TMEM are done; counters are stable
    const uint32_t g0 = bars->blocks_done, c0 = bars->chunks_done;   // running block / chunk indices of this CTA

    float acc[BN / 2];
#pragma unroll
    for (int i = 0; i < BN / 2; ++i) acc[i] = 0.f;
    const uint32_t idesc = instr_desc();
    const uint32_t smem_base = smem_u32(smem);
    const uint64_t bdesc0 = smem_desc(smem_base);         // stage 0, B hi; other stages / lo / k-steps are address offsets (>>4)

    float4 pa[4], pb[2];
    auto fetch = [&](int kb) {
        const int k0 = k_begin + kb * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + 16 * half + 4 * i;
            pa[i] = (arow < M && k < k_end) ? *reinterpret_cast<const float4 *>(A + (size_t)arow * lda + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int row = n0 + br + 32 * i, k = k0 + 4 * bc;
            pb[i] = (row < N && k < k_end) ? *reinterpret_cast<const float4 *>(Bm + (size_t)row * ldb + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto issue = [&](int kb, int kb0, bool last) {      // warps 0..2: this warp's product of block kb
        const uint32_t g = g0 + (uint32_t)kb;
        const int s = (int)(g % NSTAGE);
        mbar_wait(bars->full + s, (g / NSTAGE) & 1u, 100 + kb);
        tc_fence_after();
        if (lane == 0) {
            const uint64_t bhi = bdesc0 + (uint64_t)((s * (2 * B_BYTES)) >> 4), blo = bhi + (uint64_t)(B_BYTES >> 4);
            const uint32_t ahi = tmem_base + COL_A + s * 64, alo = ahi + 32;
            const uint32_t a_op = warp == 1 ? alo : ahi;
            const uint64_t b_op = warp == 2 ? blo : bhi;
            const uint32_t d_op = tmem_base + (warp == 0 ? COL_MAIN : warp == 1 ? COL_CORR : COL_CORR2);
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)   // 8 tf32 = 8 TMEM columns = 32 B = 2 descriptor units
                mma_tf32_ts(d_op, a_op + 8 * k, b_op + 2 * k, idesc, (kb == kb0 && k == 0) ? 0u : 1u);
            mma_commit(bars->empty + s);                 // stage free once all three products have read it
            if (last) mma_commit(&bars->chunk);
        }
        __syncwarp();
    };

    for (int ch = 0; ch < nchunks; ++ch) {
        const int kb0 = ch * KB_PER_CHUNK, kb1 = min(nkb, kb0 + KB_PER_CHUNK);
        fetch(kb0);
        for (int kb = kb0; kb < kb1; ++kb) {
            const uint32_t g = g0 + (uint32_t)kb;
            const int s = (int)(g % NSTAGE);
            float4 ca[4], cb[2];
            TC_T0
#pragma unroll
            for (int i = 0; i < 4; ++i) ca[i] = pa[i];
#pragma unroll
            for (int i = 0; i < 2; ++i) cb[i] = pb[i];
#if HIRL_TC_DEBUG == 4
            if (__float_as_uint(ca[0].x) == 0x12345678u && __float_as_uint(cb[1].x) == 0x12345678u && __float_as_uint(ca[3].x) == 0x12345678u) g_tc_dbg[63] = 1;
#endif
            TC_T(warp) TC_TR
            if (kb + 1 < kb1) fetch(kb + 1);                                              // in flight while this block is staged
            if (g >= NSTAGE) mbar_wait(bars->empty + s, (g / NSTAGE - 1u) & 1u, 200 + kb);  // MMAs that read this stage are done
            tc_fence_after();
            TC_T(8 + warp) TC_TR
            // A: my row's 16 elements -> hi / lo -> my TMEM lane
            {
                float hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float x[4] = {ca[i].x, ca[i].y, ca[i].z, ca[i].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        hi[4 * i + j] = __uint_as_float(__float_as_uint(x[j]) & 0xFFFFE000u);
                        lo[4 * i + j] = x[j] - hi[4 * i + j];
                    }
                }
                const uint32_t ta = lane_addr + COL_A + s * 64 + 16 * half;
                tmem_st16(ta, hi);
                tmem_st16(ta + 32, lo);
            }
            // B: two 16-byte chunks -> hi / lo -> swizzled shared memory
            {
                const uint32_t st = smem_base + s * (2 * B_BYTES);
#pragma unroll
                for (int i = 0; i < 2; ++i) split_store(st, st + B_BYTES, swz(br + 32 * i, bc), cb[i]);
            }
            TC_T(16 + warp) TC_TR
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            fence_async_smem();              // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars->full + s);
            TC_T(24 + warp) TC_TR
            if (warp < ISSUERS && kb > kb0) issue(kb - 1, kb0, false);   // one block behind: its operands are certainly complete
            TC_T(32 + warp)
        }
        if (warp < ISSUERS) issue(kb1 - 1, kb0, true);
        // ---------------- chunk epilogue: TMEM -> registers (all warps) ----------------
        TC_T0
        mbar_wait(&bars->chunk, (c0 + (uint32_t)ch) & 1u, 300 + ch);
        tc_fence_after();
        TC_T(40 + warp) TC_TR
#pragma unroll
        for (int j = 0; j < BN / 2 / 16; ++j) {
            const uint32_t taddr = lane_addr + (uint32_t)(half * (BN / 2) + j * 16);
            float vm[16], vc[16], vd[16];
            tmem_ld16(taddr + COL_MAIN, vm);
            tmem_ld16(taddr + COL_CORR, vc);
            tmem_ld16(taddr + COL_CORR2, vd);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[j * 16 + i] += vm[i] + (vc[i] + vd[i]);
        }
        tc_fence_before();
        __syncthreads();                     // TMEM may be overwritten by the next chunk / tile after this
        TC_T(48 + warp)
    }
    if (tid == 0) { bars->blocks_done = g0 + (uint32_t)nkb; bars->chunks_done = c0 + (uint32_t)nchunks; }   // read after the next entry sync
    const int col0 = n0 + half * (BN / 2);
    if (arow < M && col0 < N) {
        float *dst = C + (size_t)arow * ldc + col0;
#pragma unroll
        for (int i = 0; i < BN / 2; i += 4) {
            float4 o = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            if (bias) {
                const float4 bq = *reinterpret_cast<const float4 *>(bias + col0 + i);
                o.x += bq.x; o.y += bq.y; o.z += bq.z; o.w += bq.w;
            }
            *reinterpret_cast<float4 *>(dst + i) = o;
        }
    }
}

}  // namespace tc
}  // namespace hirl
