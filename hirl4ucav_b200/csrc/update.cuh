// update.cuh -- device code of the fused HIRL update (reference hirl/agents/HIRL.py:221-334).
//
// One learn() call is a fixed sequence of stages (enum Stage in common.cuh) executed by a
// "team" of CTAs that synchronise with a team barrier:
//   R stages  one warp per batch row: gather, 13/17->256 first layers, LayerNorm, heads,
//             losses and the backward passes through heads / LayerNorms (warp-shuffle
//             reductions, column sums through shared-memory tiles, per-CTA partial
//             gradients that are reduced in a fixed order -> run-to-run deterministic)
//   G stages  the 256<->512 contractions: 64x64 FP32-FMA tiles from shared memory (forward NT
//             register-staged, dX NN / dW TN through a cp.async ring), optional split-K for
//             small batches; or 128x64 tcgen05 3xTF32 tiles (tc_gemm.cuh) -- every contraction
//             at batch >= 512 and in ensembles, stage G1 alone at 128 <= batch < 512
// The same stage functions serve the single-agent grid-wide kernel (team = whole grid) and
// the ensemble kernel (team = 8 CTAs per replica).
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"

namespace hirl {

constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int TS = 68;                       // padded row stride of a 16x64 operand tile
constexpr int GEMM_SMEM_FLOATS = 3 * 2 * 64 * 16;  // three-stage ring of A and B k-tiles (gemm_tile); gemm_tile_regs needs 4*16*TS
constexpr int TILE_ROWS = WARPS;             // rows per row-stage chunk
constexpr int ROWSTAGE_SMEM_FLOATS = 4 * TILE_ROWS * H2 + 64 + TILE_ROWS * 32;
constexpr int SMEM_FLOATS = ROWSTAGE_SMEM_FLOATS > GEMM_SMEM_FLOATS ? ROWSTAGE_SMEM_FLOATS : GEMM_SMEM_FLOATS;
constexpr size_t SMEM_BYTES = SMEM_FLOATS * sizeof(float) > (size_t)tc::SMEM_BYTES ? SMEM_FLOATS * sizeof(float) : (size_t)tc::SMEM_BYTES;

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 ldcg4(const float *p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
#ifndef HIRL_TILE_LOAD_CG
#define HIRL_TILE_LOAD_CG 0
#endif
// GEMM operand tiles: plain (L1-allocating, weak) loads by default.  Coherence is provided by
// the team barrier (its acquire fence invalidates L1), so weak loads are safe and can be served
// by the near L2 slice; -DHIRL_TILE_LOAD_CG=1 forces .cg (strong, home-slice) loads.
__device__ __forceinline__ float4 ldtile4(const float *p) {
#if HIRL_TILE_LOAD_CG
    return ldcg4(p);
#else
    return ld4(p);
#endif
}
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float act_f(float pre, float slope) { return pre > 0.f ? pre : pre * slope; }
__device__ __forceinline__ float act_g(float pre, float slope) { return pre > 0.f ? 1.f : slope; }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// A team is the set of CTAs that cooperate on one agent's update.
constexpr int FLAG_STRIDE = 32;  // one 128 B line per CTA flag
struct Team {
    int rank, size;
    unsigned *flags;  // [size][FLAG_STRIDE], zeroed by the host before every launch
    unsigned gen;
    // tensor-core tile context of this CTA (TMEM allocated once per kernel)
    uint32_t tmem_base;
    char *tc_smem;        // 1024-byte aligned view of the dynamic shared memory
    tc::Barriers *tc_bars;
};

// All-to-all flag barrier: every CTA publishes its generation with one release store (no
// atomics, no contended line) and warp 0 polls everybody's flag.  The trailing fence is the
// acquire side: it also invalidates this SM's L1, so plain loads issued afterwards observe
// everything the other CTAs wrote before arriving.  (release store + fence.acq_rel instead of two
// fence.sc: 1.6 % faster on the batch-256 step, same results.)
__device__ __forceinline__ void team_sync(Team &t) {
    // Stage functions end in lane-divergent tails (e.g. `if (lane == 0) *rs_out = rstd;`): compute-sanitizer's synccheck
    // found warps arriving at the aligned barrier below before they had reconverged ("divergent thread(s) in block",
    // profiles/r2_sanitizer_summary.txt).  An aligned bar.sync in a diverged warp is undefined behaviour, so converge first.
    __syncwarp();
    __syncthreads();
    t.gen += 1u;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) st_release_u32(t.flags + t.rank * FLAG_STRIDE, t.gen);   // gpu-scope release: cumulative over the bar.sync above
        // all of this lane's flags are polled together (independent loads in flight)
        bool pending = true;
        while (pending) {
            pending = false;
#pragma unroll
            for (int i = 0; i < (MAX_CTAS + 31) / 32; ++i) {
                const int r = threadIdx.x + 32 * i;
                if (r < t.size && (int)(ld_relaxed_u32(t.flags + r * FLAG_STRIDE) - t.gen) < 0) pending = true;
            }
            pending = __any_sync(0xffffffffu, pending);
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");   // relaxed polls + fence = acquire; also invalidates this SM's L1
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// row-wise building blocks (one warp = one batch row)
// ---------------------------------------------------------------------------------------------
// first layer + LayerNorm + activation: h1 = act(LN(W1 x + b1))   (HIRL.py:61,64,131)
// lane owns outputs n = lane + 32 j.
template <class L>
__device__ __forceinline__ void l1_row(const float *tw, const float (&x)[L::k0], float slope, int lane,
                                       float *h1_out, float *xh_out, float *rs_out, float *h1_tile = nullptr) {
    // Warps issue in order: the first instruction that needs an outstanding load holds back every load behind it.
    // The LayerNorm affine is therefore requested up front -- placed where it is used (after the two reductions) it
    // was a second, fully exposed L2 round trip per row: R0 9.8 -> 5.5 us, R1 9.7 -> 5.7 us, R4 5.7 -> 2.8 us.
    float z[8], g1[8], be1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        z[j] = tw[L::B1 + lane + 32 * j];
        g1[j] = tw[L::G1 + lane + 32 * j];
        be1[j] = tw[L::BE1 + lane + 32 * j];
    }
#pragma unroll
    for (int k = 0; k < L::k0; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaf(x[k], tw[L::W1T + k * H1 + lane + 32 * j], z[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += z[j];
    const float mean = warp_sum(s) * (1.f / H1);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d = z[j] - mean;
        v = fmaf(d, d, v);
    }
    const float rstd = 1.f / sqrtf(warp_sum(v) * (1.f / H1) + LN_EPS);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = lane + 32 * j;
        const float xh = (z[j] - mean) * rstd;
        const float hv = act_f(fmaf(xh, g1[j], be1[j]), slope);
        h1_out[n] = hv;
        if (h1_tile) h1_tile[n] = hv;
        if (xh_out) xh_out[n] = xh;
    }
    if (rs_out && lane == 0) *rs_out = rstd;
}

// loads a 512-wide row, summing the split-K partial buffers in a fixed order.
// lane owns columns c(e) = 4*lane + 128*(e/4) + e%4.
__device__ __forceinline__ void load_row512(const float *row, int nsplit, size_t split_stride, int lane, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = ldcg4(row + 4 * lane + 128 * j);
        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
    for (int s = 1; s < nsplit; ++s) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 t = ldcg4(row + s * split_stride + 4 * lane + 128 * j);
            v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
        }
    }
}
__device__ __forceinline__ void store_row512(float *row, int lane, const float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) st4(row + 4 * lane + 128 * j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
}
__device__ __forceinline__ void load_vec512(const float *p, int lane, float (&v)[16]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 t = ld4(p + 4 * lane + 128 * j);
        v[4 * j + 0] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
}

// LayerNorm2 + activation + output layer of one row: z (pre-LN, bias included) -> xh, h, out
template <class L>
__device__ __forceinline__ void head_forward(const float *tw, const float (&z)[16], float slope, int lane,
                                             float (&xh)[16], float (&h)[16], float &rstd, float (&out)[L::nout]) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) s += z[e];
    const float mean = warp_sum(s) * (1.f / H2);
    float v = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const float d = z[e] - mean;
        v = fmaf(d, d, v);
    }
    rstd = 1.f / sqrtf(warp_sum(v) * (1.f / H2) + LN_EPS);
    float g[16], be[16];
    load_vec512(tw + L::G2, lane, g);
    load_vec512(tw + L::BE2, lane, be);
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        xh[e] = (z[e] - mean) * rstd;
        h[e] = act_f(fmaf(xh[e], g[e], be[e]), slope);
    }
#pragma unroll
    for (int o = 0; o < L::nout; ++o) {
        float w[16];
        load_vec512(tw + L::W3 + o * H2, lane, w);
        float p = 0.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) p = fmaf(h[e], w[e], p);
        out[o] = warp_sum(p) + tw[L::B3 + o];
    }
}

// backward of the same block: given d(out) returns dz (grad of the pre-LN row) and dy (grad
// at the LayerNorm output = grad of gamma*xh+beta); the caller turns dy/xh/h into the
// parameter gradients.
template <class L>
__device__ __forceinline__ void head_backward(const float *tw, const float (&xh)[16], float rstd, const float (&dout)[L::nout],
                                              float slope, int lane, float (&dz)[16], float (&dy)[16]) {
    float g[16], be[16];
    load_vec512(tw + L::G2, lane, g);
    load_vec512(tw + L::BE2, lane, be);
#pragma unroll
    for (int e = 0; e < 16; ++e) dy[e] = 0.f;
#pragma unroll
    for (int o = 0; o < L::nout; ++o) {
        float w[16];
        load_vec512(tw + L::W3 + o * H2, lane, w);
#pragma unroll
        for (int e = 0; e < 16; ++e) dy[e] = fmaf(dout[o], w[e], dy[e]);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        dy[e] *= act_g(fmaf(xh[e], g[e], be[e]), slope);
        dz[e] = dy[e] * g[e];  // d xhat
        s1 += dz[e];
        s2 = fmaf(dz[e], xh[e], s2);
    }
    s1 = warp_sum(s1) * (1.f / H2);
    s2 = warp_sum(s2) * (1.f / H2);
#pragma unroll
    for (int e = 0; e < 16; ++e) dz[e] = rstd * (dz[e] - s1 - xh[e] * s2);
}

// backward through activation + LayerNorm1 of one row.  lane owns n = lane + 32 j.
// in: dh (grad at h1), xh, rstd.  out: dz (grad of W1 x + b1), dy (grad at LN output).
template <class L>
__device__ __forceinline__ void ln1_backward(const float *tw, const float (&dh)[8], const float (&xh)[8], float rstd,
                                             float slope, int lane, float (&dz)[8], float (&dy)[8]) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = lane + 32 * j;
        const float g = tw[L::G1 + n];
        dy[j] = dh[j] * act_g(fmaf(xh[j], g, tw[L::BE1 + n]), slope);
        dz[j] = dy[j] * g;
        s1 += dz[j];
        s2 = fmaf(dz[j], xh[j], s2);
    }
    s1 = warp_sum(s1) * (1.f / H1);
    s2 = warp_sum(s2) * (1.f / H1);
#pragma unroll
    for (int j = 0; j < 8; ++j) dz[j] = rstd * (dz[j] - s1 - xh[j] * s2);
}

// ---------------------------------------------------------------------------------------------
// FP32 tile GEMM: C[64x64] (+split) = sum_k A . B, three operand arrangements
// ---------------------------------------------------------------------------------------------
enum { GEMM_NT = 0, GEMM_NN = 1, GEMM_TN = 2, GEMM_TC = 3 };   // GEMM_TC: tcgen05 tile, both operands K-major
struct GemmJob {
    const float *A, *B, *bias;
    float *C;
    int lda, ldb, ldc, M, N, K, mode, nsplit;
    size_t csplit;  // element stride between split-K partial outputs
};

// operand whose reduction index is contiguous (row r = output index)
__device__ __forceinline__ float4 load_kmajor(const float *P, int ld, int r0, int R, int k0, int K, int tid) {
    const int r = r0 + (tid >> 2), k = k0 + 4 * (tid & 3);
    if (r < R && k < K) return ldtile4(P + (size_t)r * ld + k);
    return make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void store_kmajor(float *T, float4 v, int tid) {
    const int r = tid >> 2, kq = 4 * (tid & 3);
    T[(kq + 0) * TS + r] = v.x;
    T[(kq + 1) * TS + r] = v.y;
    T[(kq + 2) * TS + r] = v.z;
    T[(kq + 3) * TS + r] = v.w;
}
// operand whose output index is contiguous (row = reduction index)
__device__ __forceinline__ float4 load_rmajor(const float *P, int ld, int c0, int Cn, int k0, int K, int tid) {
    const int k = k0 + (tid >> 4), c = c0 + 4 * (tid & 15);
    if (k < K && c < Cn) return ldtile4(P + (size_t)k * ld + c);
    return make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void store_rmajor(float *T, float4 v, int tid) { st4(T + (tid >> 4) * TS + 4 * (tid & 15), v); }

// Register-staged tile (global -> registers two k-tiles ahead -> transposing stores into [k][64] shared tiles).
// Kept for the forward GEMMs (GEMM_NT): their units are short (8 k-tiles after split-K) and this version has the
// smaller fixed cost per unit (tools/fp32_gemm_probe.cu: 9.5k vs 11.3k cycles for 8 k-tiles, 20.7k vs 20.3k for 16);
// the cp.async tile below wins on the long dW / dX units (stages G3, G6, G7: -1.2, -0.6, -1.4 us).
template <int MODE>
__device__ __noinline__ void gemm_tile_regs(const float *Ap, int lda, const float *Bp, int ldb, float *C,
                                       int ldc, const float *bias, int M, int N, int K, int m0, int n0, int kt_begin,
                                       int kt_end, float *smem) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float *As = smem, *Bs = smem + 2 * 16 * TS;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#define HIRL_FETCH(kt, ra, rb)                                                                                      \
    {                                                                                                               \
        const int k0_ = (kt) << 4;                                                                                  \
        ra = (MODE == GEMM_TN) ? load_rmajor(Ap, lda, m0, M, k0_, K, tid) : load_kmajor(Ap, lda, m0, M, k0_, K, tid); \
        rb = (MODE == GEMM_NT) ? load_kmajor(Bp, ldb, n0, N, k0_, K, tid) : load_rmajor(Bp, ldb, n0, N, k0_, K, tid); \
    }
#define HIRL_STASH(buf, ra, rb)                                                      \
    {                                                                                \
        if (MODE == GEMM_TN) store_rmajor(As + (buf) * 16 * TS, ra, tid);            \
        else store_kmajor(As + (buf) * 16 * TS, ra, tid);                            \
        if (MODE == GEMM_NT) store_kmajor(Bs + (buf) * 16 * TS, rb, tid);            \
        else store_rmajor(Bs + (buf) * 16 * TS, rb, tid);                            \
    }
#define HIRL_COMPUTE(buf)                                                                        \
    {                                                                                            \
        const float *a_ = As + (buf) * 16 * TS + ty * 4, *b_ = Bs + (buf) * 16 * TS + tx * 4;    \
        float4 av = ld4(a_), bv = ld4(b_);                                                       \
        _Pragma("unroll") for (int kk = 0; kk < 16; ++kk) {                                      \
            float4 an = av, bn = bv;                                                             \
            if (kk + 1 < 16) { an = ld4(a_ + (kk + 1) * TS); bn = ld4(b_ + (kk + 1) * TS); }     \
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};      \
            _Pragma("unroll") for (int i = 0; i < 4; ++i)                                        \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]); \
            av = an; bv = bn;                                                                    \
        }                                                                                        \
    }
    // register prefetch two k-tiles ahead, double-buffered shared memory: a k-tile's global loads
    // get two full compute iterations to land.
    float4 ra0, rb0, ra1, rb1;
    __syncthreads();  // previous user of this shared memory is done
    if (kt_begin < kt_end) {
        HIRL_FETCH(kt_begin, ra0, rb0)
        HIRL_STASH(0, ra0, rb0)
        if (kt_begin + 1 < kt_end) HIRL_FETCH(kt_begin + 1, ra1, rb1)
    }
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; kt += 2) {
        if (kt + 2 < kt_end) HIRL_FETCH(kt + 2, ra0, rb0)
        HIRL_COMPUTE(0)
        if (kt + 1 < kt_end) HIRL_STASH(1, ra1, rb1)
        __syncthreads();
        if (kt + 1 >= kt_end) break;
        if (kt + 3 < kt_end) HIRL_FETCH(kt + 3, ra1, rb1)
        HIRL_COMPUTE(1)
        if (kt + 2 < kt_end) HIRL_STASH(0, ra0, rb0)
        __syncthreads();
    }
#undef HIRL_FETCH
#undef HIRL_STASH
#undef HIRL_COMPUTE
    // epilogue
    const int n = n0 + tx * 4;
    if (n < N) {
        float4 bv4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) bv4 = ld4(bias + n);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + ty * 4 + i;
            if (m < M) st4(C + (size_t)m * ldc + n, make_float4(acc[i][0] + bv4.x, acc[i][1] + bv4.y, acc[i][2] + bv4.z, acc[i][3] + bv4.w));
        }
    }
}

// ---- asynchronous global -> shared copies (LDGSTS): no register staging, no transposing stores -------------------
__device__ __forceinline__ void cp_async16(float *dst, const float *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;   // src-size 0: the 16 bytes are zero-filled, nothing is read
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One 64x64 output tile over k-tiles [kt_begin, kt_end).  Deliberately a small stand-alone function with scalar
// arguments (inlined into the big stage functions the register allocator spilled the pipeline state).
//
// Operands are copied with cp.async.cg (L2-coherent, bypasses L1 like the ld.global.cg it replaces) through a
// 3-stage ring and stay in the orientation they have in global memory:
//   * "k-major" operand (reduction index contiguous: activations [row][k], weights W[n][k]):
//     shared layout [64 rows][16 k] with the 16-byte chunk index XOR-swizzled by (row>>1)&3; a thread owns the
//     strided rows q+16i and reads one float4 = 4 k-steps of a row -> conflict-free LDS.128, no transposition.
//     (The first version transposed through 4 scalar, 2-way-conflicting STS per operand and k-tile: tools/
//     fp32_gemm_probe.cu measures 1 440 -> 1 113 cycles per k-tile for the NT tile at lda = 256.)
//   * "row-major" operand (output index contiguous: [k][out]): shared layout [16 k][64]; a thread owns the
//     contiguous outputs 4q..4q+3 and reads one float4 per k-step.
// GEMM_NT: A k-major, B k-major.  GEMM_NN: A k-major, B row-major.  GEMM_TN: A row-major, B row-major.
constexpr int GEMM_STAGES = 3, GEMM_OPF = 64 * 16;   // floats per operand per stage
template <int MODE>
__device__ __noinline__ void gemm_tile(const float *Ap, int lda, const float *Bp, int ldb, float *C,
                                       int ldc, const float *bias, int M, int N, int K, int m0, int n0, int kt_begin,
                                       int kt_end, float *smem) {
    constexpr bool AKM = MODE != GEMM_TN, BKM = MODE == GEMM_NT;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // copy roles: k-major -> (row tid>>2, chunk tid&3); row-major -> (k tid>>4, 4 outputs at 4*(tid&15))
    const int kr = tid >> 2, kc = tid & 3, rk = tid >> 4, rc = 4 * (tid & 15);
    const int a_dst = AKM ? kr * 16 + 4 * (kc ^ ((kr >> 1) & 3)) : rk * 64 + rc;
    const int b_dst = BKM ? kr * 16 + 4 * (kc ^ ((kr >> 1) & 3)) : rk * 64 + rc;
    const bool a_in = AKM ? (m0 + kr < M) : (m0 + rc < M);
    const bool b_in = BKM ? (n0 + kr < N) : (n0 + rc < N);
    const float *a_src = AKM ? Ap + (size_t)(a_in ? m0 + kr : 0) * lda + 4 * kc : Ap + (a_in ? m0 + rc : 0);
    const float *b_src = BKM ? Bp + (size_t)(b_in ? n0 + kr : 0) * ldb + 4 * kc : Bp + (b_in ? n0 + rc : 0);
    auto issue = [&](int kt) {
        const int st = (kt - kt_begin) % GEMM_STAGES, k0 = kt << 4;
        float *as = smem + (2 * st) * GEMM_OPF, *bs = as + GEMM_OPF;
        if (AKM) cp_async16(as + a_dst, a_src + k0, a_in && k0 + 4 * kc < K);
        else { const bool v = a_in && k0 + rk < K; cp_async16(as + a_dst, a_src + (size_t)(v ? k0 + rk : 0) * lda, v); }
        if (BKM) cp_async16(bs + b_dst, b_src + k0, b_in && k0 + 4 * kc < K);
        else { const bool v = b_in && k0 + rk < K; cp_async16(bs + b_dst, b_src + (size_t)(v ? k0 + rk : 0) * ldb, v); }
    };
    // fragment offsets of this thread
    int a_off[4], a_sw[4], b_off[4], b_sw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int ra = ty + 16 * i, rb = tx + 16 * i;
        a_off[i] = ra * 16; a_sw[i] = (ra >> 1) & 3;
        b_off[i] = rb * 16; b_sw[i] = (rb >> 1) & 3;
    }
    __syncthreads();  // previous user of this shared memory is done
    if (kt_begin < kt_end) issue(kt_begin);
    cp_async_commit();
    if (kt_begin + 1 < kt_end) issue(kt_begin + 1);
    cp_async_commit();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
        cp_async_wait<1>();   // k-tile kt has landed (this thread's copies) ...
        __syncthreads();      // ... and everybody's; also: stage (kt+2)%3 is no longer being read
        if (kt + 2 < kt_end) issue(kt + 2);
        cp_async_commit();
        const int st = (kt - kt_begin) % GEMM_STAGES;
        const float *as = smem + (2 * st) * GEMM_OPF, *bs = as + GEMM_OPF;
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 4 k-steps per chunk
            float a[4][4], b[4][4];     // [i or j][kk]
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (AKM) {
                    const float4 t = ld4(as + a_off[i] + 4 * (c ^ a_sw[i]));
                    a[i][0] = t.x; a[i][1] = t.y; a[i][2] = t.z; a[i][3] = t.w;
                } else {   // i indexes the k-step here: one float4 = rows 4ty..4ty+3 at k = 4c+i
                    const float4 t = ld4(as + (4 * c + i) * 64 + 4 * ty);
                    a[0][i] = t.x; a[1][i] = t.y; a[2][i] = t.z; a[3][i] = t.w;
                }
                if (BKM) {
                    const float4 t = ld4(bs + b_off[i] + 4 * (c ^ b_sw[i]));
                    b[i][0] = t.x; b[i][1] = t.y; b[i][2] = t.z; b[i][3] = t.w;
                } else {
                    const float4 t = ld4(bs + (4 * c + i) * 64 + 4 * tx);
                    b[0][i] = t.x; b[1][i] = t.y; b[2][i] = t.z; b[3][i] = t.w;
                }
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
        }
    }
    // epilogue: rows ty+16i (A k-major) or 4ty+i; columns tx+16j (B k-major) or 4tx+j
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + (AKM ? ty + 16 * i : 4 * ty + i);
        if (m >= M) continue;
        if (BKM) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx + 16 * j;
                if (n < N) C[(size_t)m * ldc + n] = acc[i][j] + (bias ? bias[n] : 0.f);
            }
        } else {
            const int n = n0 + 4 * tx;
            if (n < N) {
                float4 bv4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias) bv4 = ld4(bias + n);
                st4(C + (size_t)m * ldc + n, make_float4(acc[i][0] + bv4.x, acc[i][1] + bv4.y, acc[i][2] + bv4.z, acc[i][3] + bv4.w));
            }
        }
    }
}

constexpr int MAX_JOBS = 8;
__device__ __forceinline__ void run_gemm_jobs(const GemmJob *jobs, int njobs, const Team &tm, float *smem) {
    int start[MAX_JOBS + 1];
    int total = 0;
    for (int j = 0; j < njobs; ++j) {
        start[j] = total;
        if (jobs[j].mode == GEMM_TC) total += ((jobs[j].M + tc::BM - 1) / tc::BM) * (jobs[j].N / tc::BN) * jobs[j].nsplit;
        else total += ((jobs[j].M + 63) >> 6) * ((jobs[j].N + 63) >> 6) * jobs[j].nsplit;
    }
    start[njobs] = total;
    // Fewer work units than CTAs: spread them over the whole rank range (unit u -> rank ceil(u*size/total)) instead
    // of packing them on ranks 0..total-1, so that the busy SMs are not concentrated on a few GPCs / L2 ports.
    const bool sparse = total < tm.size;
    int t_first = tm.rank, t_step = tm.size;
    if (sparse) {
        const int u = (int)(((long long)tm.rank * total) / tm.size);
        const bool mine = u < total && (int)(((long long)u * tm.size + total - 1) / total) == tm.rank;
        t_first = mine ? u : total;
        t_step = total;   // at most one unit per CTA
    }
    for (int t = t_first; t < total; t += t_step) {
        int j = 0;
        while (t >= start[j + 1]) ++j;
        const GemmJob &J = jobs[j];
        if (J.mode == GEMM_TC) {
            const int ntn = J.N / tc::BN, per = ((J.M + tc::BM - 1) / tc::BM) * ntn;
            int local = t - start[j];
            const int split = local / per;
            local -= split * per;
            const int nkb = (J.K + tc::BK - 1) / tc::BK;   // split-K in whole 32-wide blocks
            const int k0 = (int)(((long long)nkb * split) / J.nsplit) * tc::BK;
            const int k1 = min(J.K, (int)(((long long)nkb * (split + 1)) / J.nsplit) * tc::BK);
            tc::gemm_tile(J.A, J.lda, J.M, J.B, J.ldb, J.N, J.K, (local / ntn) * tc::BM, (local % ntn) * tc::BN, k0, k1,
                          J.C + (size_t)split * J.csplit, J.ldc, split == 0 ? J.bias : nullptr, tm.tmem_base, tm.tc_smem, tm.tc_bars);
            continue;
        }
        const int ntn = (J.N + 63) >> 6, ntm = (J.M + 63) >> 6;
        int local = t - start[j];
        const int split = local / (ntm * ntn);
        local -= split * ntm * ntn;
        const int tmi = local / ntn, tni = local - tmi * ntn;
        const int nk = (J.K + 15) >> 4;
        const int kb = (int)(((long long)nk * split) / J.nsplit), ke = (int)(((long long)nk * (split + 1)) / J.nsplit);
        float *Cs = J.C + (size_t)split * J.csplit;
        const float *bias = split == 0 ? J.bias : nullptr;
        if (J.mode == GEMM_NT) gemm_tile_regs<GEMM_NT>(J.A, J.lda, J.B, J.ldb, Cs, J.ldc, bias, J.M, J.N, J.K, tmi * 64, tni * 64, kb, ke, smem);
        else if (J.mode == GEMM_NN) gemm_tile<GEMM_NN>(J.A, J.lda, J.B, J.ldb, Cs, J.ldc, bias, J.M, J.N, J.K, tmi * 64, tni * 64, kb, ke, smem);
        else gemm_tile<GEMM_TN>(J.A, J.lda, J.B, J.ldb, Cs, J.ldc, bias, J.M, J.N, J.K, tmi * 64, tni * 64, kb, ke, smem);
    }
}

// ---------------------------------------------------------------------------------------------
// scratch addressing
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float *H1p(const AgentCtx &c, int pass) { return c.sc.H1 + (size_t)pass * c.B * H1; }
__device__ __forceinline__ float *XH1p(const AgentCtx &c, int slot) { return c.sc.XH1 + (size_t)slot * c.B * H1; }
__device__ __forceinline__ float *RS1p(const AgentCtx &c, int slot) { return c.sc.RS1 + (size_t)slot * c.B; }
__device__ __forceinline__ size_t z2_split_stride(const AgentCtx &c) { return (size_t)c.B * H2; }
__device__ __forceinline__ float *Z2p(const AgentCtx &c, int pass) { return c.sc.Z2 + (size_t)pass * c.zsplit * c.B * H2; }
__device__ __forceinline__ size_t dh1_split_stride(const AgentCtx &c) { return (size_t)c.B * H1; }
__device__ __forceinline__ float *DH1p(const AgentCtx &c, int slot) { return c.sc.DH1 + (size_t)slot * c.zsplit * c.B * H1; }
__device__ __forceinline__ int g1_split(const AgentCtx &c, const hirl_step_t &st) { return c.ns[st.do_actor ? NS_G1_ACTOR : NS_G1]; }
__device__ __forceinline__ float slope_of(const AgentCtx &c) { return c.activation == HIRL_ACT_LEAKY_RELU ? 0.01f : 0.f; }

// fixed-order sum of one word of the per-CTA partial block, evaluated by one full warp
// (lane l adds ranks l, l+32, ... in order, then a fixed butterfly): deterministic, and the
// loads of all ranks are in flight together instead of one L2 round trip per rank.
// parity: -1 all ranks, 0 even ranks only, 1 odd ranks only.
__device__ __forceinline__ float lp_sum(const AgentCtx &c, const Team &tm, int word, int parity = -1) {
    const int lane = threadIdx.x & 31;
    float s = 0.f;
    for (int r = lane; r < tm.size; r += 32)
        if (parity < 0 || (r & 1) == parity) s += c.sc.LP[r * LP_WORDS + word];
    return warp_sum(s);
}

// ---- fused data-parallel mode: all-reduce over NVLink peer memory inside the kernel ---------------------------------------
// Every rank writes its contribution into ITS exchange block, raises a generation flag (release at system scope) and then
// reads every rank's contribution straight out of the peers' memory (P2P loads over NVLink, L1 bypassed) and adds them in
// rank order -- all ranks compute bit-identical sums, so their weights never diverge.  No NCCL call, no extra launch.
__device__ __forceinline__ float *dpx(const AgentCtx &c, int r) { return reinterpret_cast<float *>(c.dp_peer[r]); }
__device__ __forceinline__ void dp_signal(const AgentCtx &c, int which, unsigned gen) {   // ONE thread, after the barrier that completes the data
    __threadfence_system();
    unsigned *f = reinterpret_cast<unsigned *>(dpx(c, c.dp_rank) + DPX_FLAGS) + which * 32;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(gen) : "memory");
}
__device__ __forceinline__ void dp_wait(const AgentCtx &c, int which, unsigned gen) {     // one full warp: lane r watches rank r
    const int lane = threadIdx.x & 31;
    if (lane < c.dp_world) {
        const unsigned *f = reinterpret_cast<const unsigned *>(dpx(c, lane) + DPX_FLAGS) + which * 32;
        const long long t0 = clock64();
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
            if ((int)(v - gen) >= 0) break;
            if (clock64() - t0 > 40000000000LL) {
                printf("hirl data-parallel: rank %d waited ~20 s for rank %d (flag %d, call %u)\n", c.dp_rank, lane, which, gen);
                __trap();
            }
        }
    }
    __syncwarp();
}
// sum over ranks of word `off` of the exchange blocks, rank order (one thread)
__device__ __forceinline__ float dp_sum_word(const AgentCtx &c, size_t off) {
    float s = 0.f;
    for (int r = 0; r < c.dp_world; ++r) s += __ldcv(dpx(c, r) + off);
    return s;
}

// eight consecutive batch entries of one column of a [TILE_ROWS][ld] shared tile -> 32 contiguous bytes of a
// transposed global buffer (full sectors instead of eight scattered 4-byte writes)
__device__ __forceinline__ void store_col8(float *dst, const float *tile_col, int ld) {
    st4(dst, make_float4(tile_col[0], tile_col[ld], tile_col[2 * ld], tile_col[3 * ld]));
    st4(dst + 4, make_float4(tile_col[4 * ld], tile_col[5 * ld], tile_col[6 * ld], tile_col[7 * ld]));
}

// ---------------------------------------------------------------------------------------------
// R0: gather + first layers
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void stage_r0(const AgentCtx &c, const hirl_step_t &st, int step, const Team &tm, float *smem) {
    const int B = c.B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    int passes[6];
    int np = 0;
    if (st.mode == 2) {            // behaviour cloning (BC.py:160-185): the expert-state actor pass only
        passes[np++] = P_AB;
    } else {
        passes[np++] = P_TA; passes[np++] = P_C1; passes[np++] = P_C2;
        if (st.do_actor) {
            passes[np++] = P_A;
            if (st.mode == 0) { passes[np++] = P_AB; passes[np++] = P_BC; }
        }
    }
    const int32_t *ion = c.idx_online + (size_t)step * B, *iex = c.idx_expert + (size_t)step * B, *ibc = c.idx_bc + (size_t)step * B;
    for (int task = tm.rank; task < nchunks * np; task += tm.size) {
        const int pass = passes[task / nchunks], chunk = task % nchunks;
        const int b = chunk * TILE_ROWS + warp;
        const int slot = bwd_slot(pass);
        // tensor-core path: back-propagated passes also keep h1 transposed ([k][b]) for the dW GEMM
        const bool transposed = c.use_tc && slot >= 0 && slot < 4;
        float *tile = transposed ? smem + warp * H1 : nullptr;
        if (b < B) {
            float w;
            if (pass == P_AB) {
                const float *src = c.bctab + (size_t)__ldg(ibc + b) * BCW;
                w = lane < BCW ? __ldg(src + lane) : 0.f;
                if (lane < BCW) c.sc.XB[(size_t)b * BCW + lane] = w;
            } else {
                const float *src = b < st.n_online ? c.online + (size_t)__ldg(ion + b) * ROW
                                                   : c.expert + (size_t)__ldg(iex + (b - st.n_online)) * ROW;
                w = __ldg(src + lane);
                if (pass == P_C1) c.sc.X[(size_t)b * ROW + lane] = w;
            }
            if (pass == P_C1 || pass == P_C2) {
                float x[SA];
#pragma unroll
                for (int k = 0; k < SA; ++k) x[k] = __shfl_sync(0xffffffffu, w, k);
                l1_row<CriticL>(c.critic + (pass == P_C2 ? CriticL::SIZE : 0), x, slope, lane, H1p(c, pass) + (size_t)b * H1,
                                XH1p(c, slot) + (size_t)b * H1, RS1p(c, slot) + b, tile);
            } else {
                float x[S];
                const int off = pass == P_TA ? SA : 0;  // target actor reads next_state
#pragma unroll
                for (int k = 0; k < S; ++k) x[k] = __shfl_sync(0xffffffffu, w, off + k);
                const float *net = pass == P_TA ? c.tactor : pass == P_BC ? c.bc : c.actor;
                l1_row<ActorL>(net, x, slope, lane, H1p(c, pass) + (size_t)b * H1,
                               slot >= 0 ? XH1p(c, slot) + (size_t)b * H1 : nullptr, slot >= 0 ? RS1p(c, slot) + b : nullptr, tile);
            }
        } else if (transposed) {
#pragma unroll
            for (int j = 0; j < 8; ++j) tile[lane + 32 * j] = 0.f;
        }
        if (transposed) {
            __syncthreads();
            if (tid < H1) store_col8(c.sc.H1T + ((size_t)slot * H1 + tid) * c.BT + chunk * TILE_ROWS, smem + tid, H1);
            __syncthreads();
        }
    }
}

// second-layer forward GEMM job: Z2[pass] = H1[pass] . W2^T + b2
template <class L>
__device__ __forceinline__ GemmJob fwd_job(const AgentCtx &c, int pass, const float *tw, int nsplit, bool tc) {
    GemmJob J;
    J.A = H1p(c, pass); J.lda = H1;
    J.B = tw + L::W2; J.ldb = H1;
    J.bias = tw + L::B2;
    J.C = Z2p(c, pass); J.ldc = H2;
    J.M = c.B; J.N = H2; J.K = H1;
    J.mode = tc ? GEMM_TC : GEMM_NT; J.nsplit = nsplit; J.csplit = z2_split_stride(c);   // (full tensor-core mode: the host sets every split to 1)
    return J;
}
// dH1[slot] = dZ2[pass] . W2
template <class L>
__device__ __forceinline__ GemmJob dx_job(const AgentCtx &c, int pass, const float *tw, int nsplit, const float *w2t = nullptr) {
    GemmJob J;
    J.A = Z2p(c, pass); J.lda = H2;
    J.B = tw + L::W2; J.ldb = H1;
    if (c.use_tc) { J.B = w2t; J.ldb = H2; }     // W2^T [H1][H2]: reduction index contiguous
    J.bias = nullptr;
    J.C = DH1p(c, bwd_slot(pass)); J.ldc = H1;
    J.M = c.B; J.N = H1; J.K = H2;
    J.mode = c.use_tc ? GEMM_TC : GEMM_NN; J.nsplit = c.use_tc ? 1 : nsplit; J.csplit = dh1_split_stride(c);
    return J;
}
// dW2 = dZ2[pass]^T . H1[pass]  (+ second branch)
template <class L>
__device__ __forceinline__ GemmJob dw_job(const AgentCtx &c, int pass, float *dst) {
    GemmJob J;
    J.A = Z2p(c, pass); J.lda = H2;
    J.B = H1p(c, pass); J.ldb = H1;
    J.K = c.B;
    J.bias = nullptr;
    J.C = dst; J.ldc = H1;
    J.M = H2; J.N = H1;
    J.mode = GEMM_TN; J.nsplit = c.dw_split; J.csplit = (size_t)H2 * H1;
    if (c.use_tc) {   // dW2[n][k] = sum_b dz2T[n][b] h1T[k][b]
        const int slot = bwd_slot(pass);
        J.A = c.sc.DZ2T + (size_t)slot * H2 * c.BT; J.lda = c.BT;
        J.B = c.sc.H1T + (size_t)slot * H1 * c.BT; J.ldb = c.BT;
        J.K = c.BT;
        J.mode = GEMM_TC;
    }
    return J;
}

__device__ __noinline__ void stage_g1(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    int n = 0;
    const int ns = g1_split(c, st);
    const bool tc = c.use_tc || c.tc_g1;
    if (st.mode == 2) {
        jobs[n++] = fwd_job<ActorL>(c, P_AB, c.actor, ns, tc);
        run_gemm_jobs(jobs, n, tm, smem);
        return;
    }
    jobs[n++] = fwd_job<ActorL>(c, P_TA, c.tactor, ns, tc);
    jobs[n++] = fwd_job<CriticL>(c, P_C1, c.critic, ns, tc);
    jobs[n++] = fwd_job<CriticL>(c, P_C2, c.critic + CriticL::SIZE, ns, tc);
    if (st.do_actor) {
        jobs[n++] = fwd_job<ActorL>(c, P_A, c.actor, ns, tc);
        if (st.mode == 0) {
            jobs[n++] = fwd_job<ActorL>(c, P_AB, c.actor, ns, tc);
            jobs[n++] = fwd_job<ActorL>(c, P_BC, c.bc, ns, tc);
        }
    }
    run_gemm_jobs(jobs, n, tm, smem);
}

// ---------------------------------------------------------------------------------------------
// R1: target-actor head -> a' -> first layers of the target critic; actor-side heads
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void stage_r1(const AgentCtx &c, const hirl_step_t &st, const Team &tm) {
    const int B = c.B, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    const int nkinds = st.do_actor ? (st.mode == 0 ? 4 : 2) : 1;
    for (int task = tm.rank; task < nchunks * nkinds; task += tm.size) {
        const int kind = task / nchunks;
        const int b = (task % nchunks) * TILE_ROWS + warp;
        if (b >= B) continue;
        const int pass = kind == 0 ? P_TA : kind == 1 ? P_A : kind == 2 ? P_AB : P_BC;
        const float *net = kind == 0 ? c.tactor : kind == 3 ? c.bc : c.actor;
        float z[16], xh[16], h[16], rstd, out[A];
        load_row512(Z2p(c, pass) + (size_t)b * H2, g1_split(c, st), z2_split_stride(c), lane, z);
        head_forward<ActorL>(net, z, slope, lane, xh, h, rstd, out);
        float a4[A];
#pragma unroll
        for (int j = 0; j < A; ++j) a4[j] = tanhf(out[j]);
        if (kind == 0) {
#pragma unroll
            for (int j = 0; j < A; ++j) {
                const float nz = fminf(fmaxf(st.noise[j], -c.noise_clamp), c.noise_clamp);   // HIRL.py:266
                a4[j] = fminf(fmaxf(a4[j] + nz, -1.f), 1.f);                                   // HIRL.py:267
            }
            if (lane < A) c.sc.APRIME[(size_t)b * A + lane] = a4[lane];
            const float w = c.sc.X[(size_t)b * ROW + lane];
            float x[SA];
#pragma unroll
            for (int k = 0; k < S; ++k) x[k] = __shfl_sync(0xffffffffu, w, SA + k);
#pragma unroll
            for (int j = 0; j < A; ++j) x[S + j] = a4[j];
            l1_row<CriticL>(c.tcritic, x, slope, lane, H1p(c, P_TC1) + (size_t)b * H1, nullptr, nullptr);
            l1_row<CriticL>(c.tcritic + CriticL::SIZE, x, slope, lane, H1p(c, P_TC2) + (size_t)b * H1, nullptr, nullptr);
        } else {
            float *dst = kind == 1 ? c.sc.PI : kind == 2 ? c.sc.PIB : c.sc.ABC;
            if (lane < A) dst[(size_t)b * A + lane] = a4[lane];
        }
    }
}

__device__ __noinline__ void stage_g2(const AgentCtx &c, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    jobs[0] = fwd_job<CriticL>(c, P_TC1, c.tcritic, c.ns[NS_G2], c.use_tc != 0);
    jobs[1] = fwd_job<CriticL>(c, P_TC2, c.tcritic + CriticL::SIZE, c.ns[NS_G2], c.use_tc != 0);
    run_gemm_jobs(jobs, 2, tm, smem);
}

// shared-memory carve-up of the row stages
struct RowSmem {
    float *t0, *t1, *t2, *t3;  // [TILE_ROWS][H2] column-sum tiles
    float *red;                // [64] small reductions
    float *aux;                // [TILE_ROWS][32] per-row side values
};
__device__ __forceinline__ RowSmem row_smem(float *smem) {
    RowSmem r;
    r.t0 = smem; r.t1 = smem + TILE_ROWS * H2; r.t2 = smem + 2 * TILE_ROWS * H2; r.t3 = smem + 3 * TILE_ROWS * H2;
    r.red = smem + 4 * TILE_ROWS * H2;
    r.aux = r.red + 64;
    return r;
}

// ---------------------------------------------------------------------------------------------
// R2: TD target (HIRL.py:270-274), critic heads, loss (HIRL.py:284), backward to dz2
//   CTAs are split into two groups, one per critic tower.
// ---------------------------------------------------------------------------------------------
__device__ __noinline__ void stage_r2(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    const int B = c.B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    const int gsize = tm.size >> 1, tower = tm.rank & 1, grank = tm.rank >> 1;
    RowSmem sm = row_smem(smem);
    float loss_acc = 0.f;
    if (grank < gsize) {
        const float *tw = c.critic + tower * CriticL::SIZE;
        const int pass = tower ? P_C2 : P_C1;
        float a_db2[2] = {0.f, 0.f}, a_dg2[2] = {0.f, 0.f}, a_dbe2[2] = {0.f, 0.f}, a_dw3[2] = {0.f, 0.f}, a_db3 = 0.f;
        for (int chunk = grank; chunk < nchunks; chunk += gsize) {
            const int b = chunk * TILE_ROWS + warp;
            float dz[16], t1v[16], t2v[16], t3v[16], dq = 0.f;
            if (b < B) {
                float z[16], xh[16], h[16], rstd, o1[1], o2[1];
                load_row512(Z2p(c, P_TC1) + (size_t)b * H2, c.ns[NS_G2], z2_split_stride(c), lane, z);
                head_forward<CriticL>(c.tcritic, z, slope, lane, xh, h, rstd, o1);
                load_row512(Z2p(c, P_TC2) + (size_t)b * H2, c.ns[NS_G2], z2_split_stride(c), lane, z);
                head_forward<CriticL>(c.tcritic + CriticL::SIZE, z, slope, lane, xh, h, rstd, o2);
                const float r = c.sc.X[(size_t)b * ROW + 30], d = c.sc.X[(size_t)b * ROW + 31];
                const float tq = fminf(o1[0], o2[0]);
                const float y = __fadd_rn(r, __fmul_rn(__fmul_rn(c.gamma, tq), __fsub_rn(1.f, d)));   // HIRL.py:274
                load_row512(Z2p(c, pass) + (size_t)b * H2, g1_split(c, st), z2_split_stride(c), lane, z);
                float q[1];
                head_forward<CriticL>(tw, z, slope, lane, xh, h, rstd, q);
                const float delta = q[0] - y;
                if (lane == 0) {
                    loss_acc = fmaf(delta, delta, loss_acc);
                    c.sc.Q[(size_t)tower * B + b] = q[0];
                    c.sc.TQ[(size_t)tower * B + b] = tower ? o2[0] : o1[0];
                    if (tower == 0) c.sc.Y[b] = y;
                }
                dq = (2.f / (float)c.Bg) * delta;  // mse_loss backward, mean reduction (over the global batch)
                float dout[1] = {dq}, dy[16];
                head_backward<CriticL>(tw, xh, rstd, dout, slope, lane, dz, dy);
                store_row512(Z2p(c, pass) + (size_t)b * H2, lane, dz);
#pragma unroll
                for (int e = 0; e < 16; ++e) { t1v[e] = dy[e] * xh[e]; t2v[e] = dy[e]; t3v[e] = dq * h[e]; }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) dz[e] = t1v[e] = t2v[e] = t3v[e] = 0.f;
            }
            store_row512(sm.t0 + warp * H2, lane, dz);
            store_row512(sm.t1 + warp * H2, lane, t1v);
            store_row512(sm.t2 + warp * H2, lane, t2v);
            store_row512(sm.t3 + warp * H2, lane, t3v);
            if (lane == 0) sm.aux[warp * 32] = dq;
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = tid + u * THREADS;
#pragma unroll
                for (int w = 0; w < TILE_ROWS; ++w) {
                    a_db2[u] += sm.t0[w * H2 + col];
                    a_dg2[u] += sm.t1[w * H2 + col];
                    a_dbe2[u] += sm.t2[w * H2 + col];
                    a_dw3[u] += sm.t3[w * H2 + col];
                }
                if (c.use_tc) store_col8(c.sc.DZ2T + ((size_t)tower * H2 + col) * c.BT + chunk * TILE_ROWS, sm.t0 + col, H2);
            }
            if (tid == 0)
                for (int w = 0; w < TILE_ROWS; ++w) a_db3 += sm.aux[w * 32];
            __syncthreads();
        }
        const int P = gsize < nchunks ? gsize : nchunks;
        if (grank < P) {
            float *part = c.sc.PART_C + ((size_t)grank * 2 + tower) * CriticL::SMALL;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = tid + u * THREADS;
                part[CriticL::small_index(CriticL::B2) + col] = a_db2[u];
                part[CriticL::small_index(CriticL::G2) + col] = a_dg2[u];
                part[CriticL::small_index(CriticL::BE2) + col] = a_dbe2[u];
                part[CriticL::small_index(CriticL::W3) + col] = a_dw3[u];
            }
            if (tid < 32) part[CriticL::small_index(CriticL::B3) + tid] = tid == 0 ? a_db3 : 0.f;
        }
    }
    // per-CTA loss partial (fixed order over warps)
    if (lane == 0) sm.red[warp] = loss_acc;
    __syncthreads();
    if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < WARPS; ++w) s += sm.red[w];
        c.sc.LP[tm.rank * LP_WORDS + 0] = s;
    }
    __syncthreads();
}

__device__ __noinline__ void stage_g3(const AgentCtx &c, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    // split-K over the batch: partial dW2 buffers, summed by the Adam stage (direct when not split)
    const size_t part = (size_t)H2 * H1 * c.dw_split;
    jobs[0] = dw_job<CriticL>(c, P_C1, c.dw_split > 1 ? c.sc.GW2P : c.g_critic + CriticL::W2);
    jobs[1] = dw_job<CriticL>(c, P_C2, c.dw_split > 1 ? c.sc.GW2P + part : c.g_critic + CriticL::SIZE + CriticL::W2);
    jobs[2] = dx_job<CriticL>(c, P_C1, c.critic, c.ns[NS_G3DX], c.w2t_critic);
    jobs[3] = dx_job<CriticL>(c, P_C2, c.critic + CriticL::SIZE, c.ns[NS_G3DX], c.w2t_critic + (size_t)H1 * H2);
    run_gemm_jobs(jobs, 4, tm, smem);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm1 / first-layer backward for one group of rows; accumulates per-thread column sums.
//   xsrc: row-major inputs (ld words per row), x = first L::k0 words.
// ---------------------------------------------------------------------------------------------
template <class L>
struct Ln1Acc {
    float db1 = 0.f, dg1 = 0.f, dbe1 = 0.f, dw[L::k0];
    __device__ Ln1Acc() {
#pragma unroll
        for (int k = 0; k < L::k0; ++k) dw[k] = 0.f;
    }
};
template <class L>
__device__ __forceinline__ void ln1_bwd_rows(const AgentCtx &c, const float *tw, int slot, int nsplit, const float *xsrc, int xld,
                                             int first_chunk, int chunk_step, Ln1Acc<L> &acc, const RowSmem &sm, float *gpi_out) {
    const int B = c.B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    float *t_dz = sm.t0, *t_dg = sm.t0 + TILE_ROWS * H1, *t_dbe = sm.t0 + 2 * TILE_ROWS * H1;
    for (int chunk = first_chunk; chunk < nchunks; chunk += chunk_step) {
        const int b = chunk * TILE_ROWS + warp;
        float dz[8], dy[8], xh[8];
        float xw = 0.f;
        if (b < B) {
            float dh[8];
            const float *dsrc = DH1p(c, slot) + (size_t)b * H1;
#pragma unroll
            for (int j = 0; j < 8; ++j) dh[j] = __ldcg(dsrc + lane + 32 * j);
            for (int s = 1; s < nsplit; ++s)
#pragma unroll
                for (int j = 0; j < 8; ++j) dh[j] += __ldcg(dsrc + s * dh1_split_stride(c) + lane + 32 * j);
            const float *xs = XH1p(c, slot) + (size_t)b * H1;
#pragma unroll
            for (int j = 0; j < 8; ++j) xh[j] = xs[lane + 32 * j];
            const float rstd = RS1p(c, slot)[b];
            ln1_backward<L>(tw, dh, xh, rstd, slope, lane, dz, dy);
            if (xsrc && lane < L::k0) xw = xsrc[(size_t)b * xld + lane];
            if (gpi_out) {  // gradient w.r.t. the action part of the critic input
                float g[A];
#pragma unroll
                for (int j2 = 0; j2 < A; ++j2) {
                    float p = 0.f;
#pragma unroll
                    for (int j = 0; j < 8; ++j) p = fmaf(dz[j], tw[L::W1T + (S + j2) * H1 + lane + 32 * j], p);
                    g[j2] = warp_sum(p);
                }
                if (lane < A) gpi_out[(size_t)b * A + lane] = g[lane];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) dz[j] = dy[j] = xh[j] = 0.f;
        }
        if (xsrc) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = lane + 32 * j;
                t_dz[warp * H1 + n] = dz[j];
                t_dg[warp * H1 + n] = dy[j] * xh[j];
                t_dbe[warp * H1 + n] = dy[j];
            }
            sm.aux[warp * 32 + lane] = xw;
            __syncthreads();
            if (tid < H1) {
#pragma unroll
                for (int w = 0; w < TILE_ROWS; ++w) {
                    const float dzv = t_dz[w * H1 + tid];
                    acc.db1 += dzv;
                    acc.dg1 += t_dg[w * H1 + tid];
                    acc.dbe1 += t_dbe[w * H1 + tid];
#pragma unroll
                    for (int k = 0; k < L::k0; ++k) acc.dw[k] = fmaf(dzv, sm.aux[w * 32 + k], acc.dw[k]);
                }
            }
            __syncthreads();
        }
    }
}
template <class L>
__device__ __forceinline__ void ln1_write_partial(float *part, const Ln1Acc<L> &acc) {
    const int tid = threadIdx.x;
    if (tid < H1) {
        part[L::B1 + tid] = acc.db1;
        part[L::G1 + tid] = acc.dg1;
        part[L::BE1 + tid] = acc.dbe1;
#pragma unroll
        for (int k = 0; k < L::k0; ++k) part[L::W1T + k * H1 + tid] = acc.dw[k];
    }
}

__device__ __noinline__ void stage_r3(const AgentCtx &c, const Team &tm, float *smem) {
    const int nchunks = (c.B + TILE_ROWS - 1) / TILE_ROWS;
    const int gsize = tm.size >> 1, tower = tm.rank & 1, grank = tm.rank >> 1;
    RowSmem sm = row_smem(smem);
    if (grank < gsize) {
        Ln1Acc<CriticL> acc;
        ln1_bwd_rows<CriticL>(c, c.critic + tower * CriticL::SIZE, tower, c.ns[NS_G3DX], c.sc.X, ROW, grank, gsize, acc, sm, nullptr);
        const int P = gsize < nchunks ? gsize : nchunks;
        if (grank < P) ln1_write_partial<CriticL>(c.sc.PART_C + ((size_t)grank * 2 + tower) * CriticL::SMALL, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam single-tensor arithmetic) and Polyak (HIRL.py:11-13)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void adam_elem(float g, float &w, float &m, float &v, float neg_step, float bc2_sqrt) {
    m = fmaf(0.1f, g - m, m);                                   // exp_avg.lerp_(grad, 1-beta1)
    v = fmaf(__fmul_rn(0.001f, g), g, __fmul_rn(v, 0.999f));    // mul_(beta2).addcmul_(g, g, 1-beta2)
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), bc2_sqrt), 1e-8f);
    w = __fadd_rn(w, __fdiv_rn(__fmul_rn(neg_step, m), denom));  // addcdiv_(m, denom, -step_size)
}
__device__ __forceinline__ float polyak_elem(float t, float p, float tau, float one_minus_tau) {
    return __fadd_rn(__fmul_rn(t, one_minus_tau), __fmul_rn(p, tau));
}

template <class L, int U>
__device__ __forceinline__ void adam_arena_impl(const Team &tm, int ntowers, float *w, float *gw, const float *w2parts, int nparts, float *m, float *v, const float *part,
                                                int part_towers, int nslots, float neg_step, float bc2_sqrt, float *target, bool polyak,
                                                float tau, float one_minus_tau, float *w2t, int phase,
                                                const AgentCtx *dpc, size_t dp_off, float *gw_out, const float *gw_in) {
    // gw_in: where a GEMM stage wrote dW2 directly (nparts == 0) when that is not gw itself (fused data-parallel: gw is the exchange buffer)
    // phase (data-parallel mode): 0 whole stage; 1 only the fixed-order sums of the partials -> gw (the gradient of this
    // shard, all-reduced between the two launches); 2 only Adam (+ Polyak) from gw; 3 (fused variant) Adam from the sum over the
    // ranks' exchange blocks at float offset dp_off, read over NVLink in rank order; the summed gradient is kept in gw_out
    // U elements per thread and iteration: every load of the U elements (partials, peers' blocks over NVLink, w / m / v) is issued
    // before the first one is consumed -- the loop is bound by round trips (L2, or NVLink in the fused data-parallel mode), not
    // by arithmetic.  The order of the additions inside an element is unchanged, so results are bit-identical for every U.
    const int total = ntowers * L::SIZE, stride = tm.size * THREADS;
    for (int e0 = tm.rank * THREADS + threadIdx.x; e0 < total; e0 += U * stride) {
        float g[U], wv[U], mv[U], vv[U], tv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * stride;
            g[u] = 0.f;
            if (e >= total) continue;
            const int tower = e / L::SIZE, off = e - tower * L::SIZE;
            const bool in_w2 = off >= L::W2 && off < L::B2;
            if (phase == 3) {
                for (int r = 0; r < dpc->dp_world; ++r) g[u] += __ldcv(dpx(*dpc, r) + dp_off + e);
            } else if (phase == 2) {
                g[u] = __ldcg(gw + e);
            } else if (in_w2) {
                if (nparts > 0) {   // fixed-order sum of the split-K / per-branch partial buffers
                    const float *p0 = w2parts + (size_t)tower * nparts * H2 * H1 + (off - L::W2);
                    g[u] = __ldcg(p0);
                    for (int q = 1; q < nparts; ++q) g[u] += __ldcg(p0 + (size_t)q * H2 * H1);
                } else {
                    g[u] = __ldcg((gw_in ? gw_in : gw) + e);
                }
            } else {
                const int si = L::small_index(off);
                for (int s = 0; s < nslots; ++s) g[u] += __ldcg(part + ((size_t)s * part_towers + tower) * L::SMALL + si);
            }
            if (phase != 1) {
                wv[u] = w[e]; mv[u] = m[e]; vv[u] = v[e];
                tv[u] = polyak ? target[e] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int e = e0 + u * stride;
            if (e >= total) continue;
            const int tower = e / L::SIZE, off = e - tower * L::SIZE;
            const bool in_w2 = off >= L::W2 && off < L::B2;
            if (phase == 3) gw_out[e] = g[u];
            else if (phase != 2 && (!in_w2 || nparts > 0 || gw_in)) gw[e] = g[u];
            if (phase == 1) continue;
            adam_elem(g[u], wv[u], mv[u], vv[u], neg_step, bc2_sqrt);
            w[e] = wv[u]; m[e] = mv[u]; v[e] = vv[u];
            if (w2t && in_w2) {   // keep W2^T current for the tensor-core dX GEMMs
                const int i = off - L::W2, n = i / H1, k = i - n * H1;
                w2t[(size_t)tower * H1 * H2 + (size_t)k * H2 + n] = wv[u];
            }
            if (polyak) target[e] = polyak_elem(tv[u], wv[u], tau, one_minus_tau);
        }
    }
}
// one element per thread and iteration (the form round 1 shipped): kept for the many-part case, see adam_arena below
template <class L>
__device__ __forceinline__ void adam_arena_scalar(const Team &tm, int ntowers, float *w, float *gw, const float *w2parts, int nparts, float *m, float *v, const float *part,
                                                  int part_towers, int nslots, float neg_step, float bc2_sqrt, float *target, bool polyak,
                                                  float tau, float one_minus_tau, float *w2t, int phase, const float *gw_in) {
    const int total = ntowers * L::SIZE;
    for (int e = tm.rank * THREADS + threadIdx.x; e < total; e += tm.size * THREADS) {
        const int tower = e / L::SIZE, off = e - tower * L::SIZE;
        float g;
        const bool in_w2 = off >= L::W2 && off < L::B2;
        if (phase == 2) {
            g = __ldcg(gw + e);
        } else if (in_w2) {
            if (nparts > 0) {   // fixed-order sum of the split-K / per-branch partial buffers
                const float *p0 = w2parts + (size_t)tower * nparts * H2 * H1 + (off - L::W2);
                g = __ldcg(p0);
                for (int q = 1; q < nparts; ++q) g += __ldcg(p0 + (size_t)q * H2 * H1);
                gw[e] = g;
            } else {
                g = __ldcg((gw_in ? gw_in : gw) + e);
                if (gw_in) gw[e] = g;
            }
        } else {
            const int si = L::small_index(off);
            g = 0.f;
            for (int s = 0; s < nslots; ++s) g += __ldcg(part + ((size_t)s * part_towers + tower) * L::SMALL + si);
            gw[e] = g;
        }
        if (phase == 1) continue;
        float wv = w[e], mv = m[e], vv = v[e];
        adam_elem(g, wv, mv, vv, neg_step, bc2_sqrt);
        w[e] = wv; m[e] = mv; v[e] = vv;
        if (w2t && in_w2) {   // keep W2^T current for the tensor-core dX GEMMs
            const int i = off - L::W2, n = i / H1, k = i - n * H1;
            w2t[(size_t)tower * H1 * H2 + (size_t)k * H2 + n] = wv;
        }
        if (polyak) target[e] = polyak_elem(target[e], wv, tau, one_minus_tau);
    }
}
// U = 4 where an element has few loads of its own (ensembles and small batches: dW2 in <= 2 parts; the peers' blocks of the fused
// data-parallel all-reduce): ensemble128 +7 %.  With the 8-16 split-K parts of batch 4096 the batched form is SLOWER (stage G4
// 22 -> 30 us, tools/stage_profile.py), so that case keeps one element per iteration.
template <class L>
__device__ __forceinline__ void adam_arena(const Team &tm, int ntowers, float *w, float *gw, const float *w2parts, int nparts, float *m, float *v, const float *part,
                                           int part_towers, int nslots, float neg_step, float bc2_sqrt, float *target, bool polyak,
                                           float tau, float one_minus_tau, float *w2t, int phase = 0,
                                           const AgentCtx *dpc = nullptr, size_t dp_off = 0, float *gw_out = nullptr, const float *gw_in = nullptr) {
    if (phase == 3 || nparts <= 2)
        adam_arena_impl<L, 4>(tm, ntowers, w, gw, w2parts, nparts, m, v, part, part_towers, nslots, neg_step, bc2_sqrt, target, polyak, tau, one_minus_tau, w2t, phase,
                              dpc, dp_off, gw_out, gw_in);
    else
        adam_arena_scalar<L>(tm, ntowers, w, gw, w2parts, nparts, m, v, part, part_towers, nslots, neg_step, bc2_sqrt, target, polyak, tau, one_minus_tau, w2t, phase, gw_in);
}
template <class L>
__device__ __forceinline__ void polyak_arena(const Team &tm, int ntowers, const float *w, float *target, float tau, float one_minus_tau) {
    const int total = ntowers * L::SIZE;
    for (int e = tm.rank * THREADS + threadIdx.x; e < total; e += tm.size * THREADS)
        target[e] = polyak_elem(target[e], w[e], tau, one_minus_tau);
}

// In the Adam stages the last rank also reduces the loss scalars (a few dependent L2 round trips): with an equal
// share of the arena it was the stage's straggler by 1.3-1.7 us (tools/stage_profile.py --ranks).  On large teams
// the arena is therefore dealt to ranks 0..size-2 only and the last rank just does the reductions.
__device__ __forceinline__ Team arena_team(const Team &tm) {
    Team t = tm;
    if (tm.size >= 32) t.size = tm.size - 1;
    return t;
}

__device__ __noinline__ void stage_g4(const AgentCtx &c, const hirl_step_t &st, Team &tm, int phase, unsigned gen) {
    const int nchunks = (c.B + TILE_ROWS - 1) / TILE_ROWS, gsize = tm.size >> 1;
    const int P = gsize < nchunks ? gsize : nchunks;
    const Team at = arena_team(tm);
    const bool fused = phase == 4;      // fused data-parallel variant: shard sums -> exchange block -> in-kernel all-reduce -> Adam
    const size_t par = gen & 1u;
    float *gsum = fused ? dpx(c, c.dp_rank) + DPX_GC + par * CRITIC_SIZE : c.g_critic;
    if (tm.rank < at.size)
        adam_arena<CriticL>(at, 2, c.critic, gsum, c.sc.GW2P, c.dw_split > 1 ? c.dw_split : 0, c.m_critic, c.v_critic, c.sc.PART_C, 2, P,
                            st.critic_neg_step, st.critic_bc2_sqrt, nullptr, false, 0.f, 0.f, c.use_tc ? c.w2t_critic : nullptr, fused ? 1 : phase,
                            nullptr, 0, nullptr, fused ? c.g_critic : nullptr);
    if (phase != 2 && tm.rank == tm.size - 1 && threadIdx.x < 32) {
        // critic_loss = mse(Q1,y) + mse(Q2,y): even ranks hold tower 0, odd ranks tower 1
        const float s0 = lp_sum(c, tm, 0, 0), s1 = lp_sum(c, tm, 0, 1);
        if (threadIdx.x == 0) {
            const float loss = __fadd_rn(s0 / (float)c.Bg, s1 / (float)c.Bg);
            if (fused) dpx(c, c.dp_rank)[DPX_SC + par * 32 + 0] = loss;
            else c.carry[0] = loss;
        }
    }
    if (fused) {
        team_sync(tm);                                                 // this rank's gradient and loss are complete
        if (tm.rank == 0 && threadIdx.x == 0) dp_signal(c, DPF_CRITIC, gen);
        if (threadIdx.x < 32) dp_wait(c, DPF_CRITIC, gen);
        __syncthreads();
        if (tm.rank < at.size)
            adam_arena<CriticL>(at, 2, c.critic, gsum, c.sc.GW2P, 0, c.m_critic, c.v_critic, c.sc.PART_C, 2, P, st.critic_neg_step, st.critic_bc2_sqrt,
                                nullptr, false, 0.f, 0.f, c.use_tc ? c.w2t_critic : nullptr, 3, &c, DPX_GC + par * CRITIC_SIZE, c.g_critic);
        if (tm.rank == tm.size - 1 && threadIdx.x == 0) c.carry[0] = dp_sum_word(c, DPX_SC + par * 32 + 0);
    }
}

// ---------------------------------------------------------------------------------------------
// actor phase
// ---------------------------------------------------------------------------------------------
// R4: first layers of Q1(s, pi(s)) and Q1(s, a_bc) with the updated critic (HIRL.py:296, 302)
__device__ __noinline__ void stage_r4(const AgentCtx &c, const hirl_step_t &st, const Team &tm) {
    const int B = c.B, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    const bool soft = st.mode == 0 && st.bc_weight_now == 100.f;
    const int nkinds = soft ? 2 : 1;
    for (int task = tm.rank; task < nchunks * nkinds; task += tm.size) {
        const int kind = task / nchunks;
        const int b = (task % nchunks) * TILE_ROWS + warp;
        if (b >= B) continue;
        const float w = c.sc.X[(size_t)b * ROW + lane];
        const float *act = (kind == 0 ? c.sc.PI : c.sc.ABC) + (size_t)b * A;
        float x[SA];
#pragma unroll
        for (int k = 0; k < S; ++k) x[k] = __shfl_sync(0xffffffffu, w, k);
#pragma unroll
        for (int j = 0; j < A; ++j) x[S + j] = act[j];
        if (kind == 0)
            l1_row<CriticL>(c.critic, x, slope, lane, H1p(c, P_QP) + (size_t)b * H1, XH1p(c, 4) + (size_t)b * H1, RS1p(c, 4) + b);
        else
            l1_row<CriticL>(c.critic, x, slope, lane, H1p(c, P_QS) + (size_t)b * H1, nullptr, nullptr);
    }
}
__device__ __noinline__ void stage_g5(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    int n = 0;
    jobs[n++] = fwd_job<CriticL>(c, P_QP, c.critic, c.ns[NS_G5], c.use_tc != 0);
    if (st.mode == 0 && st.bc_weight_now == 100.f) jobs[n++] = fwd_job<CriticL>(c, P_QS, c.critic, c.ns[NS_G5], c.use_tc != 0);
    run_gemm_jobs(jobs, n, tm, smem);
}
// R5: Q heads, soft count (HIRL.py:303), rl_loss partial (HIRL.py:297), backward of Q1(s,pi) with
// the unit upstream gradient -1/B; the (1-w) factor is applied where it meets the actor (R6).
__device__ __noinline__ void stage_r5(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    const int B = c.B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    const bool soft = st.mode == 0 && st.bc_weight_now == 100.f;
    RowSmem sm = row_smem(smem);
    float cnt = 0.f, sum_q = 0.f;
    for (int chunk = tm.rank; chunk < nchunks; chunk += tm.size) {
        const int b = chunk * TILE_ROWS + warp;
        if (b >= B) continue;
        float z[16], xh[16], h[16], rstd, q[1];
        float softq = 0.f;
        if (soft) {
            float rs2, qs[1];
            load_row512(Z2p(c, P_QS) + (size_t)b * H2, c.ns[NS_G5], z2_split_stride(c), lane, z);
            head_forward<CriticL>(c.critic, z, slope, lane, xh, h, rs2, qs);
            softq = qs[0];
        }
        load_row512(Z2p(c, P_QP) + (size_t)b * H2, c.ns[NS_G5], z2_split_stride(c), lane, z);
        head_forward<CriticL>(c.critic, z, slope, lane, xh, h, rstd, q);
        if (lane == 0) {
            c.sc.RLQ[b] = q[0];
            sum_q += q[0];
            if (soft) {
                c.sc.SOFTQ[b] = softq;
                cnt += softq > q[0] ? 1.f : 0.f;
            }
        }
        float dout[1] = {-1.f / (float)c.Bg}, dz[16], dy[16];
        head_backward<CriticL>(c.critic, xh, rstd, dout, slope, lane, dz, dy);
        store_row512(Z2p(c, P_QP) + (size_t)b * H2, lane, dz);
    }
    if (lane == 0) { sm.red[warp] = cnt; sm.red[8 + warp] = sum_q; }
    __syncthreads();
    if (tid == 0) {
        float s0 = 0.f, s1 = 0.f;
        for (int w = 0; w < WARPS; ++w) { s0 += sm.red[w]; s1 += sm.red[8 + w]; }
        c.sc.LP[tm.rank * LP_WORDS + 1] = s0;
        c.sc.LP[tm.rank * LP_WORDS + 2] = s1;
    }
    __syncthreads();
}
__device__ __noinline__ void stage_g6(const AgentCtx &c, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    jobs[0] = dx_job<CriticL>(c, P_QP, c.critic, c.ns[NS_G6], c.w2t_critic);
    run_gemm_jobs(jobs, 1, tm, smem);
}

// the behaviour-cloning weight of this call (HIRL.py:292-308), evaluated identically by every CTA
// (call with a full warp: the soft count is a warp-collective sum)
struct BcWeight { float w, one_minus_w, count; };
__device__ __forceinline__ BcWeight bc_weight_of(const AgentCtx &c, const hirl_step_t &st, const Team &tm, unsigned gen = 0u) {
    BcWeight r;
    r.count = 0.f;
    double w = (double)st.bc_weight_now;
    if (st.mode == 2) w = 1.0;       // BC: the loss IS the cloning term
    else if (st.mode != 0) w = 0.0;
    else if (st.bc_weight_now == 100.f) {
        if (c.dp_world > 0) {   // fused data-parallel: every rank published its shard's count before stage G6 (learn_step); sum in rank order
            dp_wait(c, DPF_COUNT, gen);
            const int lane = threadIdx.x & 31;
            const float mine = lane < c.dp_world ? __ldcv(dpx(c, lane) + DPX_SC + (gen & 1u) * 32 + 5) : 0.f;
            float cnt = 0.f;
            for (int q = 0; q < c.dp_world; ++q) cnt += __shfl_sync(0xffffffffu, mine, q);
            r.count = cnt;
        } else
        r.count = c.dp_count ? __ldcg(c.dp_count) : lp_sum(c, tm, 1);   // NCCL data-parallel: the all-reduced count of every shard
        const float mean = st.force_soft_weight >= 0.f ? st.force_soft_weight : r.count / (float)c.Bg;  // .float().mean()
        w = (double)mean + (double)st.bc_warm_up;
    }
    if (w > 1.0) w = 1.0;
    r.w = (float)w;
    r.one_minus_w = (float)(1.0 - w);
    return r;
}

// R6: group 0: d pi from the critic, actor head backward on batch states (rl branch);
//     group 1: BC loss (HIRL.py:310-311) and actor head backward on expert states.
__device__ __noinline__ void stage_r6(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem, unsigned gen) {
    const int B = c.B, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunks = (B + TILE_ROWS - 1) / TILE_ROWS;
    const float slope = slope_of(c);
    const int gsize = tm.size >> 1, branch = tm.rank & 1, grank = tm.rank >> 1;
    RowSmem sm = row_smem(smem);
    if (tid < 32) {
        const BcWeight b0 = bc_weight_of(c, st, tm, gen);   // warp-collective
        if (tid == 0) { sm.red[32] = b0.w; sm.red[33] = b0.one_minus_w; }
    }
    __syncthreads();
    BcWeight bw;
    bw.w = sm.red[32]; bw.one_minus_w = sm.red[33]; bw.count = 0.f;
    float bc_sq = 0.f, fire_sq = 0.f;
    const bool active = grank < gsize && (branch == 0 ? st.mode != 2 : st.mode != 1);
    const float bc_scale = st.mode == 2 ? 1.f : bw.w * c.loss_lambda;   // BC.py:179 has no loss_lambda
    if (active) {
        // ---- step 1 (branch 0): LayerNorm1 backward of Q1(s,pi) -> unit gradient w.r.t. pi
        if (branch == 0) {
            Ln1Acc<CriticL> unused;
            ln1_bwd_rows<CriticL>(c, c.critic, 4, c.ns[NS_G6], nullptr, 0, grank, gsize, unused, sm, c.sc.GPI);
            __syncwarp();
        }
        const int pass = branch ? P_AB : P_A;
        float a_db2[2] = {0.f, 0.f}, a_dg2[2] = {0.f, 0.f}, a_dbe2[2] = {0.f, 0.f}, a_dw3[A][2], a_db3 = 0.f;
#pragma unroll
        for (int j = 0; j < A; ++j) a_dw3[j][0] = a_dw3[j][1] = 0.f;
        for (int chunk = grank; chunk < nchunks; chunk += gsize) {
            const int b = chunk * TILE_ROWS + warp;
            float dz[16], t1v[16], t2v[16], hv[16], dpre[A] = {0.f, 0.f, 0.f, 0.f};
            if (b < B) {
                float z[16], xh[16], rstd, out[A];
                load_row512(Z2p(c, pass) + (size_t)b * H2, c.ns[NS_G1_ACTOR], z2_split_stride(c), lane, z);
                head_forward<ActorL>(c.actor, z, slope, lane, xh, hv, rstd, out);
#pragma unroll
                for (int j = 0; j < A; ++j) {
                    const float pi = tanhf(out[j]);
                    float dpi;
                    if (branch == 0) {
                        dpi = bw.one_minus_w * c.sc.GPI[(size_t)b * A + j];
                    } else {
                        const float diff = pi - c.sc.XB[(size_t)b * BCW + S + j];
                        if (lane == 0) {
                            bc_sq = fmaf(diff, diff, bc_sq);
                            if (j == 3) fire_sq = fmaf(diff, diff, fire_sq);
                        }
                        dpi = ((2.f / (float)(A * c.Bg)) * diff) * bc_scale;
                    }
                    dpre[j] = dpi * (1.f - pi * pi);   // tanh backward
                }
                float dy[16];
                head_backward<ActorL>(c.actor, xh, rstd, dpre, slope, lane, dz, dy);
                store_row512(Z2p(c, pass) + (size_t)b * H2, lane, dz);
#pragma unroll
                for (int e = 0; e < 16; ++e) { t1v[e] = dy[e] * xh[e]; t2v[e] = dy[e]; }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) dz[e] = t1v[e] = t2v[e] = hv[e] = 0.f;
            }
            store_row512(sm.t0 + warp * H2, lane, dz);
            store_row512(sm.t1 + warp * H2, lane, t1v);
            store_row512(sm.t2 + warp * H2, lane, t2v);
            store_row512(sm.t3 + warp * H2, lane, hv);
            if (lane < A) sm.aux[warp * 32 + lane] = dpre[lane];
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = tid + u * THREADS;
#pragma unroll
                for (int w = 0; w < TILE_ROWS; ++w) {
                    a_db2[u] += sm.t0[w * H2 + col];
                    a_dg2[u] += sm.t1[w * H2 + col];
                    a_dbe2[u] += sm.t2[w * H2 + col];
                    const float hh = sm.t3[w * H2 + col];
#pragma unroll
                    for (int j = 0; j < A; ++j) a_dw3[j][u] = fmaf(sm.aux[w * 32 + j], hh, a_dw3[j][u]);
                }
                if (c.use_tc) store_col8(c.sc.DZ2T + ((size_t)(2 + branch) * H2 + col) * c.BT + chunk * TILE_ROWS, sm.t0 + col, H2);
            }
            if (tid < A)
                for (int w = 0; w < TILE_ROWS; ++w) a_db3 += sm.aux[w * 32 + tid];
            __syncthreads();
        }
        const int P = gsize < nchunks ? gsize : nchunks;
        if (grank < P) {
            float *part = c.sc.PART_A + ((size_t)branch * P + grank) * ActorL::SMALL;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = tid + u * THREADS;
                part[ActorL::small_index(ActorL::B2) + col] = a_db2[u];
                part[ActorL::small_index(ActorL::G2) + col] = a_dg2[u];
                part[ActorL::small_index(ActorL::BE2) + col] = a_dbe2[u];
#pragma unroll
                for (int j = 0; j < A; ++j) part[ActorL::small_index(ActorL::W3) + j * H2 + col] = a_dw3[j][u];
            }
            if (tid < 32) part[ActorL::small_index(ActorL::B3) + tid] = tid < A ? a_db3 : 0.f;
        }
    }
    if (lane == 0) { sm.red[warp] = bc_sq; sm.red[8 + warp] = fire_sq; }
    __syncthreads();
    if (tid == 0) {
        float s0 = 0.f, s1 = 0.f;
        for (int w = 0; w < WARPS; ++w) { s0 += sm.red[w]; s1 += sm.red[8 + w]; }
        c.sc.LP[tm.rank * LP_WORDS + 3] = s0;
        c.sc.LP[tm.rank * LP_WORDS + 4] = s1;
    }
    __syncthreads();
}
__device__ __noinline__ void stage_g7(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    GemmJob jobs[MAX_JOBS];
    int n = 0;
    // the two branches' dW2 (each split-K over the batch) go to separate partial buffers; Adam adds them
    if (st.mode != 2) jobs[n++] = dw_job<ActorL>(c, P_A, c.sc.GW2P);
    if (st.mode != 1) jobs[n++] = dw_job<ActorL>(c, P_AB, c.sc.GW2P + (size_t)H2 * H1 * c.dw_split);
    if (st.mode != 2) jobs[n++] = dx_job<ActorL>(c, P_A, c.actor, c.ns[NS_G7DX], c.w2t_actor);
    if (st.mode != 1) jobs[n++] = dx_job<ActorL>(c, P_AB, c.actor, c.ns[NS_G7DX], c.w2t_actor);
    run_gemm_jobs(jobs, n, tm, smem);
}
__device__ __noinline__ void stage_r7(const AgentCtx &c, const hirl_step_t &st, const Team &tm, float *smem) {
    const int nchunks = (c.B + TILE_ROWS - 1) / TILE_ROWS;
    const int gsize = tm.size >> 1, branch = tm.rank & 1, grank = tm.rank >> 1;
    RowSmem sm = row_smem(smem);
    if (grank < gsize && (branch == 0 ? st.mode != 2 : st.mode != 1)) {
        Ln1Acc<ActorL> acc;
        ln1_bwd_rows<ActorL>(c, c.actor, branch ? 3 : 2, c.ns[NS_G7DX], branch ? c.sc.XB : c.sc.X, branch ? BCW : ROW, grank, gsize, acc, sm, nullptr);
        const int P = gsize < nchunks ? gsize : nchunks;
        if (grank < P) ln1_write_partial<ActorL>(c.sc.PART_A + ((size_t)branch * P + grank) * ActorL::SMALL, acc);
    }
}
// forceinline, unlike the other stages.  Round 1: built as a separate function this stage once read a wrong dW2-partials base in
// warps 1-7 (compute-sanitizer: out-of-bounds __ldcg in adam_arena) and any probe made the fault disappear.  Round 2
// (tools/sanitize_all.sh, profiles/r2_sanitizer_summary_*.txt): rebuilt as a separate function (HIRL_G8_NOINLINE=1) the stage is
// clean under memcheck and initcheck, and the only hazard racecheck found in the kernel -- the SHARED step record `st` being
// rewritten by a warp that is already in the next call while a slower warp still reads it -- is exactly a mechanism that hands
// this stage a stale st.mode / st.do_polyak (its partials base and slot count derive from them); it is fixed (learn_step takes
// register copies, stages run between barriers) and the fault has not reproduced since.  The inlined form stays: one call site.
#ifndef HIRL_G8_NOINLINE
#define HIRL_G8_NOINLINE 0   // 1: build stage_g8 as a separate function again (tools/sanitize_all.sh re-tests the round-1 fault)
#endif
#if HIRL_G8_NOINLINE
__device__ __noinline__ void stage_g8(const AgentCtx &c, const hirl_step_t &st, Team &tm, int phase, unsigned gen) {
#else
__device__ __forceinline__ void stage_g8(const AgentCtx &c, const hirl_step_t &st, Team &tm, int phase, unsigned gen) {
#endif
    const int nchunks = (c.B + TILE_ROWS - 1) / TILE_ROWS, gsize = tm.size >> 1;
    const int P = gsize < nchunks ? gsize : nchunks;
    const int nslots = st.mode == 0 ? 2 * P : P;
    const float tau = c.tau, omt = (float)(1.0 - (double)c.tau);
    // mode 2 (BC) only has the expert-state branch: its partial slots / dW2 parts are the second half
    const size_t part_off = st.mode == 2 ? (size_t)P * ActorL::SMALL : 0, w2_off = st.mode == 2 ? (size_t)H2 * H1 * c.dw_split : 0;
    const float *part_a = c.sc.PART_A + part_off;
    const float *w2parts = c.sc.GW2P + w2_off;
    const Team at = arena_team(tm);
    const bool fused = phase == 4;      // fused data-parallel variant (see stage_g4)
    const size_t par = gen & 1u;
    float *gsum = fused ? dpx(c, c.dp_rank) + DPX_GA + par * ACTOR_SIZE : c.g_actor;
    if (tm.rank < at.size)
        adam_arena<ActorL>(at, 1, c.actor, gsum, w2parts, (st.mode == 0 ? 2 : 1) * c.dw_split, c.m_actor, c.v_actor, part_a, 1, nslots, st.actor_neg_step, st.actor_bc2_sqrt,
                       c.tactor, st.do_polyak != 0, tau, omt, c.use_tc ? c.w2t_actor : nullptr, fused ? 1 : phase, nullptr, 0, nullptr, fused ? c.g_actor : nullptr);
    if (!fused && phase != 1 && st.do_polyak && tm.rank < at.size) polyak_arena<CriticL>(at, 2, c.critic, c.tcritic, tau, omt);
    if (phase != 2 && tm.rank == tm.size - 1 && threadIdx.x < 32) {
        const BcWeight bw = bc_weight_of(c, st, tm, gen);
        const float rl_loss = st.mode == 2 ? 0.f : -(lp_sum(c, tm, 2) / (float)c.Bg);     // HIRL.py:297
        float bc_loss = 0.f, fire = 0.f;
        if (st.mode == 0) {
            bc_loss = __fmul_rn(lp_sum(c, tm, 3) / (float)(A * c.Bg), c.loss_lambda);      // HIRL.py:311
            fire = lp_sum(c, tm, 4) / (float)c.Bg;                                         // HIRL.py:319 (x lambda on the host, in double)
        } else if (st.mode == 2) {
            bc_loss = lp_sum(c, tm, 3) / (float)(A * c.Bg);                                // BC.py:179
        }
        if (threadIdx.x == 0) {
            // (fused data-parallel: these are this shard's parts of the global means; they are summed over the ranks below)
            float *dst = fused ? dpx(c, c.dp_rank) + DPX_SC + par * 32 : c.carry;
            dst[1] = __fadd_rn(__fmul_rn(bc_loss, bw.w), __fmul_rn(rl_loss, bw.one_minus_w));  // HIRL.py:321
            dst[2] = bc_loss;
            dst[3] = rl_loss;
            dst[4] = fire;
            c.carry[5] = bw.w;
            c.carry[6] = bw.count;
        }
    }
    if (fused) {
        team_sync(tm);                                                 // this rank's gradient and scalars are complete
        if (tm.rank == 0 && threadIdx.x == 0) dp_signal(c, DPF_ACTOR, gen);
        if (threadIdx.x < 32) dp_wait(c, DPF_ACTOR, gen);
        __syncthreads();
        if (tm.rank < at.size) {
            adam_arena<ActorL>(at, 1, c.actor, gsum, w2parts, 0, c.m_actor, c.v_actor, part_a, 1, nslots, st.actor_neg_step, st.actor_bc2_sqrt,
                               c.tactor, st.do_polyak != 0, tau, omt, c.use_tc ? c.w2t_actor : nullptr, 3, &c, DPX_GA + par * ACTOR_SIZE, c.g_actor);
            if (st.do_polyak) polyak_arena<CriticL>(at, 2, c.critic, c.tcritic, tau, omt);
        }
        if (tm.rank == tm.size - 1 && threadIdx.x == 0) {
#pragma unroll
            for (int j = 1; j <= 4; ++j) c.carry[j] = dp_sum_word(c, DPX_SC + par * 32 + j);
        }
    }
}

// one full learn() call for one agent; returns true when the debug stop point was reached
// first / last / adam_phase: the data-parallel mode (hirl_allreduce_init) runs ONE call as several launches over stage
// ranges [first, last] with the gradient all-reduces between them; every stage keeps its state in global memory, so a
// range simply resumes where the previous launch stopped.  Defaults: whole call (first = 1, last = ST_END, phase 0).
__device__ __forceinline__ bool learn_step(const AgentCtx &c, int step, Team &tm, float *smem, int stop_after, long long *prof,
                                           int first = 1, int last = ST_END, int adam_phase = 0, unsigned dp_base = 0u) {
    // adam_phase 4: fused data-parallel variant (in-kernel all-reduce over NVLink peer memory); gen: this call's number
    const unsigned gen = dp_base + (unsigned)step + 1u;
    __shared__ hirl_step_t st;
    if (threadIdx.x < (int)(sizeof(hirl_step_t) / sizeof(int)))
        reinterpret_cast<int *>(&st)[threadIdx.x] = reinterpret_cast<const int *>(c.steps + step)[threadIdx.x];
    __syncthreads();
    if (prof && step >= PROF_STEPS) prof = nullptr;
    if (prof) {
        prof += ((size_t)step * ST_END * MAX_CTAS + tm.rank) * 2;
        if (threadIdx.x == 0) prof[1] = clock64();   // stage 0 slot: step start
    }
#define HIRL_STAGE(id, call)                                                            \
    if ((id) >= first) {                                                                \
    call;                                                                               \
    if (prof && threadIdx.x == 0) prof[(size_t)(id) * MAX_CTAS * 2] = clock64();        \
    team_sync(tm);                                                                      \
    if (prof && threadIdx.x == 0) prof[(size_t)(id) * MAX_CTAS * 2 + 1] = clock64();    \
    if (stop_after == (id)) return true;                                                \
    if (last == (id)) {                                                                 \
        if ((id) == ST_G6 && c.dp_count && tm.rank == tm.size - 1 && threadIdx.x < 32) { \
            const float cnt_ = lp_sum(c, tm, 1);   /* this shard's count of soft_Q > rl_Q, all-reduced before stage R6 */ \
            if (threadIdx.x == 0) *c.dp_count = cnt_;                                   \
        }                                                                               \
        const bool end_ = adam_phase != 1 && ((id) == ST_G8 || ((id) == ST_G4 && !do_actor)); \
        if (end_ && tm.rank == tm.size - 1 && threadIdx.x < HIRL_OUT_SCALARS)          \
            c.out_scalars[(size_t)step * HIRL_OUT_SCALARS + threadIdx.x] = c.carry[threadIdx.x]; \
        return true;                                                                    \
    }                                                                                   \
    }
    // NOTE: every stage function has exactly ONE call site; modes skip stages with conditions instead.  (Round 1 saw a
    // corrupted pointer in the actor Adam with a second call site; see the note at stage_g8 for what round 2's sanitizer sweep
    // found -- the shared step record race below -- which explains that symptom as well.)
    // Control flow uses register copies taken here: racecheck flagged the read of st.do_actor AFTER the last team barrier of
    // a critic-only call against the next call's write of `st` by a warp that is already one call ahead.
    const bool bc_only = st.mode == 2;   // behaviour cloning (BC.py:160-185): expert-state actor pass, backward, Adam
    const bool do_actor = st.do_actor != 0;
    HIRL_STAGE(ST_R0, stage_r0(c, st, step, tm, smem))
    HIRL_STAGE(ST_G1, stage_g1(c, st, tm, smem))
    if (!bc_only) {
        HIRL_STAGE(ST_R1, stage_r1(c, st, tm))
        HIRL_STAGE(ST_G2, stage_g2(c, tm, smem))
        HIRL_STAGE(ST_R2, stage_r2(c, st, tm, smem))
        HIRL_STAGE(ST_G3, stage_g3(c, tm, smem))
        HIRL_STAGE(ST_R3, stage_r3(c, tm, smem))
        HIRL_STAGE(ST_G4, stage_g4(c, st, tm, adam_phase, gen))
    }
    if (bc_only || do_actor) {
        if (!bc_only) {
            HIRL_STAGE(ST_R4, stage_r4(c, st, tm))
            HIRL_STAGE(ST_G5, stage_g5(c, st, tm, smem))
            HIRL_STAGE(ST_R5, stage_r5(c, st, tm, smem))
            if (adam_phase == 4 && tm.rank == tm.size - 1 && threadIdx.x < 32) {
                // fused data-parallel: this shard's count of soft_Q > rl_Q (complete since stage R5's barrier) goes out now;
                // the peers' counts travel over NVLink while stage G6 computes and are summed at the top of stage R6
                const float cnt_ = lp_sum(c, tm, 1);
                if (threadIdx.x == 0) {
                    dpx(c, c.dp_rank)[DPX_SC + (gen & 1u) * 32 + 5] = cnt_;
                    dp_signal(c, DPF_COUNT, gen);
                }
            }
            HIRL_STAGE(ST_G6, stage_g6(c, tm, smem))
        }
        HIRL_STAGE(ST_R6, stage_r6(c, st, tm, smem, gen))
        HIRL_STAGE(ST_G7, stage_g7(c, st, tm, smem))
        HIRL_STAGE(ST_R7, stage_r7(c, st, tm, smem))
        HIRL_STAGE(ST_G8, stage_g8(c, st, tm, adam_phase, gen))
    }
#undef HIRL_STAGE
    if (tm.rank == tm.size - 1 && threadIdx.x < HIRL_OUT_SCALARS) c.out_scalars[(size_t)step * HIRL_OUT_SCALARS + threadIdx.x] = c.carry[threadIdx.x];
    return false;
}

}  // namespace hirl
