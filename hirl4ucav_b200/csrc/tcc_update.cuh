// tcc_update.cuh -- the tensor-core cluster kernel: the fused HIRL update (reference hirl/agents/HIRL.py:221-334) for a
// single agent at batch <= 256, re-cut so that a thread-block CLUSTER of 8 CTAs carries a block of 64 batch rows through
// whole networks (forward AND backward) with every 256<->512 contraction on tcgen05 (3xTF32), and grid-wide barriers
// remain only where a weight gradient has to be complete.
//
//   phase A  units, one cluster each (12 at batch 256; B200 co-schedules 15 clusters of 8 CTAs):
//              target-actor unit r : gather rows -> target actor on s' -> a' published through L2 (flag)
//              tower unit (r, t)   : gather rows -> critic tower t forward on (s, a) [overlaps the target actor]
//                                    -> target tower t on (s', a') -> tq_t published to the partner tower's unit (flag)
//                                    -> TD target y -> loss -> backward: dz2, dH1 (K-split over the cluster, summed
//                                    through L2), LayerNorm1 / first-layer backward, dW2 (per row block) -> partials
//   ---- grid barrier ----
//   phase B  critic Adam over the whole grid (fixed-order sums of the partials); refreshes the tensor-core operand
//            images of the weights it changed
//   ---- grid barrier ----                                                                  (critic-only calls end here)
//   phase C  units (r, role): role A : actor(s) -> pi -> Q1(s,pi) [updated critic] -> backward to d pi -> actor backward
//                             role S : bc_actor(s) -> a_bc -> Q1(s,a_bc) -> soft Q published to role A (flag)
//                             role AB: actor(s_bc) -> pi_bc -> cloning loss -> actor backward
//            Both actor branches are linear in their loss scale, so they are back-propagated UNSCALED and the Q-filter
//            weight w (HIRL.py:303-308; needs every row block's count of soft_Q > rl_Q) is applied to the summed partial
//            gradients of a branch in phase D: (1 - w) resp. w * lambda.  Nobody waits for the count.
//   ---- grid barrier ----
//   phase D  actor Adam, Polyak every 3rd actor update (HIRL.py:327-330), operand images, returned scalars
//   ---- grid barrier, only if the targets moved or the next call is an actor call (phase A of a critic-only call reads
//        nothing else that phase D writes) ----
//
// CTA j of a cluster owns the output features 64 j .. 64 j + 63 of the 512-wide layers.  Per pass:
//   * first layer (K = 13 / 17) + LayerNorm1 + activation ONCE PER CLUSTER: warp w of CTA j evaluates batch row 8 j + w in
//     FP32 FMAs (weights staged in shared memory by one bulk copy) and writes it hi/lo-split into the cluster's H1 image
//     in L2, which has the layout of the SWIZZLE_128B K-major operand tile H1[64][256] (128 KB); after one cluster
//     barrier every CTA multicasts one k-block (hi + lo, 16 KB) of the image into the operand tile of all 8 CTAs
//     (cp.async.bulk ... .multicast::cluster).  For back-propagated passes the CTA also writes its 8 rows of the
//     TRANSPOSED tile image (tf32 has no MN-major twin of this layout, tools/tcc_probe.cu);
//   * second layer: tcgen05.mma M=64 N=64, A = H1 tile, B = this CTA's slice of a pre-split, pre-swizzled image of W2
//     streamed by cp.async.bulk through a 3-slot ring (the first three units are requested at the end of the PREVIOUS
//     pass, as soon as its MMAs have released the ring); the three products of the 3xTF32 split are issued by three threads
//     into three TMEM accumulators (a single issuing thread costs ~90 cycles per MMA and was the critical path);
//   * LayerNorm2 / head from the accumulators with ONE THREAD PER ROW (tcgen05.ld): row statistics and head dot products
//     are completed across the 8 CTAs by st.async pushes into every CTA's shared memory, each signalling the receiver's
//     mbarrier -- no shuffles, no cluster barrier, no global round trip;
//   * backward: dz2 per thread-row -> K-major tile;  dH1^T[256][64] = W2T-image . dz2  (M=128 x 2, this CTA's 64
//     features = its K slice) -> partial through L2 -> row owner sums 8 partials -> LayerNorm1 backward;
//     dW2^T[256][64] = H1^T . dz2^T with the H1^T image bulk-loaded from L2 and dz2^T written into the idle ring; the
//     dW2 MMAs run while the dH1 partials are exchanged.
// Everything is reduced in a fixed order: results are run-to-run deterministic.  FP32-accurate: rel-L2 <= 1e-5 against
// torch fp32 per step (tests/test_gpu_update.py, back-end "tcc").
#pragma once
#include "tcc_mma.cuh"
#include "update.cuh"

namespace hirl {
namespace tcc {

constexpr int CL = 8;              // CTAs per cluster
constexpr int RB = 64;             // batch rows per row block (= MMA M)
constexpr int NSL = H2 / CL;       // 64 output features per CTA
constexpr int TTHREADS = 256;
constexpr int MAX_NCL = 16;        // clusters per launch
constexpr int MAX_RB = 4;          // row blocks (batch <= 256)
constexpr int SLOT_BYTES = 16384, NSLOT = 3;
constexpr int ISSUERS = 3;         // one thread per product of the 3xTF32 split
// shared-memory carve-up (bytes from a 1024-aligned base)
constexpr int O_R = 0;                                   // 128 KB  H1 [8 blocks][64 x 128 B] hi | lo   /  H1^T [2 blocks][256 x 128 B] hi | lo  / staging
constexpr int R_LO = 65536;
constexpr int O_RING = 131072;                           // 48 KB   weight stream ring  /  dz2^T [2 blocks][64 x 128 B] hi | lo
constexpr int O_D = O_RING + NSLOT * SLOT_BYTES;         // 33 KB   [0,16K) exchange slots, [16K,33K) W1 of the pass | [0,32K) dz2 tile, first-layer column sums
constexpr int O_W1 = O_D + 16384;
constexpr int O_X = O_D + 16384 + 17408;                        // 8 KB    gathered rows [64][32]
constexpr int O_ACT = O_X + 8192;                        // 1 KB    action columns of the critic input [64][4]
constexpr int O_GPI = O_ACT + 1024;                      // 1 KB    d pi [64][4]
constexpr int O_PC = O_GPI + 1024;                       // 2 KB    parameter cache of the pass: b2 | g2 | be2 | w3[nout] (this CTA's 64 features) | b3
constexpr int O_RED = O_PC + 2048;                       // 2 KB    half-row combine scratch of the exchanges, one KB per exchange parity
constexpr int SMEM_USED = O_RED + 2048;
constexpr size_t SMEM_BYTES = SMEM_USED + 1024;          // + alignment slack
static_assert(SMEM_BYTES + 1200 <= 232448, "shared memory budget (227 KB per CTA incl. static)");
// TMEM columns: forward accumulators (main, Alo.Bhi, Ahi.Blo); backward accumulators (m-block, product) at 64 (3 mb + p)
constexpr int T_MAIN = 0, T_C1 = 64, T_C2 = 128;

// image geometry (floats)
constexpr size_t IMG_FWD = (size_t)H2 * H1 * 2;          // per tower: [8 slices][8 k-blocks][hi 2048 | lo 2048]
constexpr size_t IMG_T = (size_t)H2 * H1 * 2;            // per tower: [8 slices][2 m-blocks][2 k-chunks][hi 4096 | lo 4096]
constexpr size_t H1T_FLOATS = 32768;                     // H1^T image of one row block: hi [2 blocks][256 x 32] | lo
__host__ __device__ __forceinline__ uint32_t sw128_words(int row, int kk) {
    return (uint32_t)((row >> 3) * 256 + (row & 7) * 32 + ((((kk >> 2) ^ (row & 7)) << 2) | (kk & 3)));
}
__host__ __device__ __forceinline__ size_t img_fwd_index(int n, int k) {     // W2[n][k] -> hi word (lo at + 2048)
    return ((size_t)(n >> 6) * 8 + (k >> 5)) * 4096 + sw128_words(n & 63, k & 31);
}
__host__ __device__ __forceinline__ size_t img_t_index(int n, int k) {       // W2[n][k] -> hi word (lo at + 4096)
    return ((((size_t)(n >> 6) * 2 + (k >> 7)) * 2 + ((n & 63) >> 5)) * 2) * 4096 + sw128_words(k & 127, n & 31);
}

// p2p flag words (each on its own 128-byte line)
enum { PF_TQ = 0 /* [r][t] */, PF_RLQ = 8 /* [r] */, PF_AP = 12 /* [r] */, PF_WORDS = 16 };
constexpr int PROF_WORDS = 96;
enum { B_FULL = 0, B_EMPTY = NSLOT, B_ACC = 2 * NSLOT, B_W1 = 2 * NSLOT + 2, B_H1T, B_H1, B_X0, B_X1, N_BARS };

struct TccLaunch {
    AgentCtx agent;
    unsigned *bar_flags;   // [MAX_NCL][32] cluster flags of the grid barrier
    int k_steps, ncl;
    long long *prof;       // [PROF_STEPS][PROF_WORDS] SM-clock stamps of cluster 0 / CTA 0 (may be null)
    unsigned jitter;       // debug (HIRL_TCC_JITTER=seed): pseudo-random per-CTA delays at the synchronisation sites; results must not change
    unsigned jitter_ns;    // upper bound of one delay (HIRL_TCC_JITTER_NS, default 4000)
};

struct Ctl {
    char *sm;
    unsigned pad;      // sm = (dynamic shared memory symbol) + pad
    uint64_t *bars;
    uint32_t tm;
    int crank, cid, ncl, warp, lane, sub, wh, row;   // row: batch row of this thread inside the block (thread-per-row epilogues)
    bool act;                                         // lane < 16: this thread owns (row, 32 columns)
    unsigned ring_n, acc_n[2], w1_n, h1t_n, h1_n, xr, gen, pgen;
    float *h1img;      // this cluster's H1 image in L2 (layout of the shared operand tile)
    const float *pre;  // weight stream whose first NSLOT units are already on their way into the ring (issued at the end of the previous pass)
    unsigned *flags, *pflags;
    long long *pr, *pr2;   // pr: stamps of cluster 0 / CTA 0;  pr2: of the first online unit's cluster (phase A) / CTA 0
    int pb;   // base index of the fine-grained stamps of the current pass (0: off)
    unsigned jit, jit_n, jit_ns;   // timing-perturbation test: seed (0 = off), site counter, delay bound
};
__device__ __forceinline__ void stamp(const Ctl &k, int id) { if (k.pr && threadIdx.x == 0) k.pr[id] = clock64(); }
__device__ __forceinline__ void stamp2(const Ctl &k, int id) { if (k.pr2 && threadIdx.x == 0) k.pr2[id] = clock64(); }
__device__ __forceinline__ void pstamp(const Ctl &k, int i) {
    long long *p = k.pr ? k.pr : k.pr2;
    if (p && k.pb && threadIdx.x == 0) p[k.pb + i] = clock64();
}

// ---- synchronisation -----------------------------------------------------------------------------------------------
// Timing-perturbation hook: with HIRL_TCC_JITTER set every CTA (and, at warp-level sites, every warp) sleeps a pseudo-random
// 0-4 us (HIRL_TCC_JITTER_NS: other bound) at the synchronisation sites.  The kernel's results must be bit-identical with and without it
// (tests/test_gpu_update.py::test_tcc_results_do_not_depend_on_timing): every cross-CTA hand-over is ordered by a barrier,
// an mbarrier or a release/acquire flag, never by timing.
__device__ __noinline__ void jitter_sleep(unsigned seed, unsigned site, unsigned who, unsigned bound_ns) {
    unsigned h = seed * 2654435761u ^ (site * 40503u) ^ (who * 2246822519u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    if ((h & 3u) == 0u) return;
    for (unsigned left = h % bound_ns; left > 0u;) {   // __nanosleep is capped near 1 ms per call
        const unsigned d = left > 500000u ? 500000u : left;
        __nanosleep(d);
        left -= d;
    }
}
__device__ __forceinline__ void jitter(Ctl &k, bool per_warp = false) {
    if (k.jit == 0u) return;
    k.jit_n += 1u;
    jitter_sleep(k.jit, k.jit_n, (unsigned)(k.cid * CL + k.crank) * 8u + (per_warp ? (unsigned)k.warp : 0u), k.jit_ns);
}
__device__ __forceinline__ void grid_sync(Ctl &k) {
    jitter(k);
    cluster_sync();                                                    // the whole cluster has arrived (cluster-scope release / acquire)
    k.gen += 1u;
    if (threadIdx.x < 32) {                                            // one flag per cluster, polled by every CTA (no second cluster barrier)
        if (k.crank == 0 && threadIdx.x == 0) st_release_u32(k.flags + k.cid * FLAG_STRIDE, k.gen);
        bool pending = true;
        const long long t0 = clock64();
        while (pending) {
            pending = (int)threadIdx.x < k.ncl && (int)(ld_relaxed_u32(k.flags + threadIdx.x * FLAG_STRIDE) - k.gen) < 0;
            pending = __any_sync(0xffffffffu, pending);
            if (pending && clock64() - t0 > 8000000000LL) {
                if (threadIdx.x == 0) printf("hirl tcc kernel: grid barrier timed out (cluster %d gen %u)\n", k.cid, k.gen);
                __trap();
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}
// producer side of a point-to-point flag: the CTA's global writes, then one release store
__device__ __forceinline__ void flag_publish(Ctl &k, int word, unsigned value) {
    jitter(k);
    __syncthreads();
    if (threadIdx.x == 0) st_release_u32(k.pflags + word * FLAG_STRIDE, value);
}
__device__ __forceinline__ void flag_wait(Ctl &k, int word, unsigned value) {
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while ((int)(ld_acquire_u32(k.pflags + word * FLAG_STRIDE) - value) < 0) {
            if (clock64() - t0 > 8000000000LL) {
                printf("hirl tcc kernel: flag %d timed out (cluster %d cta %d want %u)\n", word, k.cid, k.crank, value);
                __trap();
            }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

// ---- cluster-wide per-row exchange ------------------------------------------------------------------------------------
// Every CTA contributes one float4 per row (its two column halves combined locally first); the contribution is pushed
// into the slot [src CTA][row] of EVERY CTA with st.async, which signals the receiver's mbarrier (16 bytes each); a round
// is complete for a CTA when its own mbarrier has seen 8 x 64 x 16 bytes.  Two slot buffers / barriers alternate.
// `chan`: the halves are (mean, M2) pairs over 32 values each and are merged with Chan's formula, else plain sums.
__device__ __forceinline__ float4 *xslot(const Ctl &k, int buf, int src, int row) {
    return reinterpret_cast<float4 *>(k.sm + O_D + buf * 8192 + (src * RB + row) * 16);
}
__device__ __forceinline__ void exchange_push(Ctl &k, float a, float b, float c, float d, bool chan) {
    const int buf = k.xr & 1;
    float4 *half = reinterpret_cast<float4 *>(k.sm + O_RED + 1024 * buf);         // [64 rows] contribution of the upper column half (one buffer per parity:
                                                                                  //  the next write is behind the next exchange's bar.sync, not only behind the mbarrier)
    if (k.act && k.wh == 1) half[k.row] = make_float4(a, b, c, d);
    if (threadIdx.x == 0) mbar_expect_tx(k.bars + B_X0 + buf, CL * RB * 16);
    __syncthreads();
    jitter(k, true);
    if (k.act && k.wh == 0) {
        const float4 o = half[k.row];
        float4 v;
        if (chan) {
            const float mean = 0.5f * (a + o.x), dl = o.x - a;
            v = make_float4(mean, (b + o.y) + 16.f * dl * dl, 0.f, 0.f);       // n = 32 + 32: M2 = M2a + M2b + d^2 * 32*32/64
        } else v = make_float4(a + o.x, b + o.y, c + o.z, d + o.w);
        const uint32_t local = smem_u32(xslot(k, buf, k.crank, k.row)), bar = smem_u32(k.bars + B_X0 + buf);
#pragma unroll
        for (int dst = 0; dst < CL; ++dst) st_async_f32x4(map_to_cta(local, dst), v.x, v.y, v.z, v.w, map_to_cta(bar, dst));
    }
    tc::mbar_wait(k.bars + B_X0 + buf, (k.xr >> 1) & 1u, 800);
    k.xr += 1u;
}
__device__ __forceinline__ float4 exchange_get(const Ctl &k, int src) {   // after exchange_push: contribution of CTA `src` for this thread's row
    return *xslot(k, (k.xr - 1u) & 1, src, k.row);
}

// ---- weight stream ------------------------------------------------------------------------------------------------
// units of 16 KB, contiguous in global memory.  Called by ALL threads (uniform arguments); lane 0 of warp 3 issues the copies.
constexpr int PRODUCER = 96;
__device__ __forceinline__ void stream_units(Ctl &k, const float *src, int first, int count) {
    if (threadIdx.x == PRODUCER) {
        for (int u = first; u < first + count; ++u) {
            const unsigned n = k.ring_n + (unsigned)u;
            const int s = (int)(n % NSLOT);
            if (n >= NSLOT) tc::mbar_wait(k.bars + B_EMPTY + s, ((n / NSLOT) - 1u) & 1u, 500 + u);
            mbar_expect_tx(k.bars + B_FULL + s, SLOT_BYTES);
            bulk_g2s(k.sm + O_RING + s * SLOT_BYTES, reinterpret_cast<const char *>(src) + (size_t)u * SLOT_BYTES, SLOT_BYTES, k.bars + B_FULL + s);
        }
    }
}
// first units of a pass: skipped if the previous pass already requested exactly this stream
__device__ __forceinline__ void stream_begin(Ctl &k, const float *units) {
    if (k.pre != units) stream_units(k, units, 0, NSLOT);
    k.pre = nullptr;
}
// the ring is idle (every MMA that read it has completed): request the next pass' first units now, so that their L2 round
// trip overlaps this pass' epilogue and the next pass' prologue
__device__ __forceinline__ void stream_prefetch(Ctl &k, const float *next_units) {
    if (next_units) stream_units(k, next_units, 0, NSLOT);
    k.pre = next_units;
}
__device__ __forceinline__ const float *slice_of(const Ctl &k, const float *img) { return img + (size_t)k.crank * 8 * 4096; }
__device__ __forceinline__ void ring_wait_full(Ctl &k, unsigned n) {   // issuer threads
    tc::mbar_wait(k.bars + B_FULL + (n % NSLOT), (n / NSLOT) & 1u, 600);
    tc::tc_fence_after();
}
__device__ __forceinline__ void acc_wait(Ctl &k, int which) {          // all threads
    tc::mbar_wait(k.bars + B_ACC + which, k.acc_n[which] & 1u, 700 + which);
    k.acc_n[which] += 1u;
    tc::tc_fence_after();
}
// operands written with ordinary stores become visible to the tensor core (async proxy); then a CTA barrier
__device__ __forceinline__ void publish_operands() {
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
}

// ---- first layer + LayerNorm1 + activation of the CTA's OWN 8 rows -> H1 image of the row block in L2 ------------------
// Warp w evaluates batch row 8 crank + w (lane owns n = lane + 32 j); the result goes hi/lo-split into the cluster's H1
// image in global memory, which has exactly the layout of the shared operand tile (8 blocks of [64 rows][32] hi | lo), so
// that after one cluster barrier every CTA multicasts one k-block of it (hi + lo, 16 KB) into the operand tile of all
// eight CTAs: the first layer is computed once per cluster instead of once per CTA.
// x_k of row b:  k < 13: xs[b][soff + k];  k >= 13 (critic): acts[b][k - 13]  (acts == nullptr: xs[b][k], the stored action)
// W1T [k0 x 256] is staged in shared memory (stage_w1); b1 / g1 / be1 come from the arena.
// own: xhat1 / rstd1 of the row, kept for the backward pass.  h1t (may be null): global H1^T image; this warp writes its column.
struct OwnRow { float xh1[8], rstd1; };
template <class L>
__device__ __forceinline__ void l1_rows(Ctl &k, const float *tw, int soff, const float *acts, float slope, float *h1img, float (&hrow)[8], OwnRow &own) {
    const float *xs = reinterpret_cast<const float *>(k.sm + O_X), *w1s = reinterpret_cast<const float *>(k.sm + O_W1);
    const int lane = k.lane, b = 8 * k.crank + k.warp;
    float b1[8], g1[8], be1[8], z[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {                                      // in flight while the bulk copy of W1 lands
        b1[j] = tw[L::B1 + lane + 32 * j];
        g1[j] = tw[L::G1 + lane + 32 * j];
        be1[j] = tw[L::BE1 + lane + 32 * j];
    }
    float xv = 0.f;
    if (lane < L::k0) xv = lane < S ? xs[b * 32 + soff + lane] : (acts ? acts[b * 4 + (lane - S)] : xs[b * 32 + lane]);
    tc::mbar_wait(k.bars + B_W1, k.w1_n & 1u, 900);
    k.w1_n += 1u;
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = b1[j];
#pragma unroll
    for (int kk = 0; kk < L::k0; ++kk) {
        const float x = __shfl_sync(0xffffffffu, xv, kk);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaf(x, w1s[kk * H1 + lane + 32 * j], z[j]);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += z[j];
    s = warp_sum(s) * (1.f / H1);
    float v = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float d = z[j] - s;
        v = fmaf(d, d, v);
    }
    v = warp_sum(v);
    const float rstd = 1.f / sqrtf(v * (1.f / H1) + LN_EPS);
    own.rstd1 = rstd;
    float *dst = h1img + sw128_words(b, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float xh = (z[j] - s) * rstd;
        const float hv = act_f(fmaf(xh, g1[j], be1[j]), slope);
        const float hi = tf32_hi(hv);
        own.xh1[j] = xh;
        dst[j * 2048] = hi;
        dst[16384 + j * 2048] = hv - hi;
        hrow[j] = hv;
    }
}
// H1^T image [256 n][64 b] (K-major along b) of a back-propagated pass: written AFTER the multicast of the forward tile has
// been issued and WHILE the MMAs are issued (warps 0-2 issue, warp 3 streams the weights, warps 4-7 write the image): off the
// critical path.  Staged through the W1 region, which is dead once every warp has left the first layer (the cluster barrier
// in between).  A writer thread handles two columns n: this CTA's 8 batch rows = one 32-byte sector hi, one lo each.
// Visibility for the bulk load in pass_backward: the next cluster barrier / grid barrier of the unit (see the call sites).
__device__ __forceinline__ void h1t_stage(Ctl &k, const float (&hrow)[8]) {   // all warps
    float *tst = reinterpret_cast<float *>(k.sm + O_W1);               // [8 rows][256]
#pragma unroll
    for (int j = 0; j < 8; ++j) tst[k.warp * H1 + k.lane + 32 * j] = hrow[j];
    if (k.warp < 4) asm volatile("bar.arrive 1, 256;" ::: "memory");
    else asm volatile("bar.sync 1, 256;" ::: "memory");
}
__device__ __forceinline__ void h1t_write(Ctl &k, float *h1t) {               // warps 4-7
    const float *tst = reinterpret_cast<const float *>(k.sm + O_W1);
    const int b0 = 8 * k.crank;
    float *t = h1t + (size_t)(b0 >> 5) * 8192;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int n = (int)threadIdx.x - 128 + 128 * half;
        float hv[8], hi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            hv[i] = tst[i * H1 + n];
            hi[i] = tf32_hi(hv[i]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float *w = t + sw128_words(n, (b0 & 31) + 4 * q);
            st4(w, make_float4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]));
            st4(w + 16384, make_float4(hv[4 * q] - hi[4 * q], hv[4 * q + 1] - hi[4 * q + 1], hv[4 * q + 2] - hi[4 * q + 2], hv[4 * q + 3] - hi[4 * q + 3]));
        }
    }
}

// parameter cache of a pass: b2 | g2 | be2 | w3[o] restricted to this CTA's 64 features, then b3[o] at 448
template <class L>
__device__ __forceinline__ void load_pc(const Ctl &k, const float *tw) {
    float *pc = reinterpret_cast<float *>(k.sm + O_PC);
    for (int i = threadIdx.x; i < (3 + L::nout) * NSL; i += TTHREADS) {
        const int a = i >> 6, c = i & 63;
        const int off = a == 0 ? L::B2 : a == 1 ? L::G2 : a == 2 ? L::BE2 : L::W3 + (a - 3) * H2;
        pc[i] = tw[off + NSL * k.crank + c];
    }
    if (threadIdx.x < L::nout) pc[448 + threadIdx.x] = tw[L::B3 + threadIdx.x];
}
// the same in two halves: the loads are issued at the top of a pass, the shared-memory stores follow the first layer (whose own
// parameter loads are then in flight at the same time: one exposed L2 round trip instead of two)
template <class L>
__device__ __forceinline__ void pc_fetch(const Ctl &k, const float *tw, float (&v)[2], float &b3v) {
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = threadIdx.x + q * TTHREADS;
        const int a = i >> 6, c = i & 63;
        const int off = a == 0 ? L::B2 : a == 1 ? L::G2 : a == 2 ? L::BE2 : L::W3 + (a - 3) * H2;
        v[q] = i < (3 + L::nout) * NSL ? tw[off + NSL * k.crank + c] : 0.f;
    }
    b3v = threadIdx.x < L::nout ? tw[L::B3 + threadIdx.x] : 0.f;
}
template <class L>
__device__ __forceinline__ void pc_store(const Ctl &k, const float (&v)[2], float b3v) {
    float *pc = reinterpret_cast<float *>(k.sm + O_PC);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int i = threadIdx.x + q * TTHREADS;
        if (i < (3 + L::nout) * NSL) pc[i] = v[q];
    }
    if (threadIdx.x < L::nout) pc[448 + threadIdx.x] = b3v;
}
// first-layer weight matrix of the pass into shared memory (one bulk copy)
template <class L>
__device__ __forceinline__ void stage_w1(Ctl &k, const float *tw) {
    __syncthreads();                                                   // previous readers of the region (and of the exchange slots) are done
    if (threadIdx.x == PRODUCER) {
        constexpr uint32_t bytes = L::k0 * H1 * 4;
        mbar_expect_tx(k.bars + B_W1, bytes);
        bulk_g2s(k.sm + O_W1, tw + L::W1T, bytes, k.bars + B_W1);
    }
}

// ---- forward pass: first layer, second layer on the tensor cores, LayerNorm2 + head per thread-row --------------------
// out[o]: the head output of this thread's row (identical in every CTA); xh / rstd: this thread's 32 normalised columns.
template <class L>
__device__ __forceinline__ void pass_forward(Ctl &k, const float *tw, const float *img, int soff, const float *acts, float slope, float *h1t,
                                             float (&xh)[32], float &rstd, float (&out)[L::nout], OwnRow &own, const float *next_units) {
    char *sm = k.sm;
    const float *unit0 = slice_of(k, img);                             // this CTA's slice: 8 k-blocks of hi | lo
    jitter(k);
    float pcv[2], b3v;
    pc_fetch<L>(k, tw, pcv, b3v);
    stage_w1<L>(k, tw);
    stream_begin(k, unit0);                                            // lands while the first layer runs
    if (threadIdx.x == PRODUCER) mbar_expect_tx(k.bars + B_H1, 131072);
    pstamp(k, 0);
    float hrow[8];
    l1_rows<L>(k, tw, soff, acts, slope, k.h1img, hrow, own);
    pc_store<L>(k, pcv, b3v);                                          // read after the cluster barrier below
    pstamp(k, 1);
    jitter(k, true);
    // every CTA's rows of the H1 image (and of the H1^T image) are in L2, and every CTA is past its last use of the
    // operand-tile region (it arrives here only after that): the multicast below may land everywhere
    cluster_sync();
    if (k.pr && k.pb == 20 && threadIdx.x == 0) k.pr[55] = clock64();
    if (threadIdx.x == PRODUCER) {
        fence_proxy_async_all();
#ifdef TCC_H1_UNICAST
#pragma unroll
        for (int q = 0; q < 4; ++q) bulk_g2s(sm + O_R + q * 32768, k.h1img + (size_t)q * 8192, 32768, k.bars + B_H1);
#else
        bulk_g2s_mc(sm + O_R + k.crank * 8192, k.h1img + (size_t)k.crank * 2048, 8192, k.bars + B_H1, (uint16_t)0xFF);
        bulk_g2s_mc(sm + O_R + R_LO + k.crank * 8192, k.h1img + 16384 + (size_t)k.crank * 2048, 8192, k.bars + B_H1, (uint16_t)0xFF);
#endif
    }
    if (h1t) h1t_stage(k, hrow);
    stream_units(k, unit0, NSLOT, 8 - NSLOT);
    if (h1t && k.warp >= 4) h1t_write(k, h1t);                          // while warps 0-2 issue the MMAs
    if (k.lane == 0 && k.warp < ISSUERS) {
        tc::tc_fence_after();
        const int p = k.warp;
        const uint32_t id = idesc(RB, NSL, 0, 0), r = smem_u32(sm + O_R) + (p == 1 ? R_LO : 0), ring = smem_u32(sm + O_RING) + (p == 2 ? 8192 : 0);
        const uint32_t acc = k.tm + 64 * p;
        tc::mbar_wait(k.bars + B_H1, k.h1_n & 1u, 960);                  // the H1 tile has landed (async proxy: no generic-proxy fence needed)
        if (k.pr && k.pb == 20 && p == 0) k.pr[56] = clock64();
        for (int c = 0; c < 8; ++c) {
            const unsigned n = k.ring_n + (unsigned)c;
            ring_wait_full(k, n);
            if (k.pr && k.pb == 20 && p == 0) k.pr[57 + c] = clock64();
            const uint64_t a = desc_k(r + c * 8192), b = desc_k(ring + (n % NSLOT) * SLOT_BYTES);
#pragma unroll
            for (int s = 0; s < 4; ++s) mma_ss(acc, a + 2 * s, b + 2 * s, id, (c == 0 && s == 0) ? 0u : 1u);
            tc::mma_commit(k.bars + B_EMPTY + (n % NSLOT));
        }
        tc::mma_commit(k.bars + B_ACC + 0);
    }
    __syncwarp();
    k.ring_n += 8u;
    k.h1_n += 1u;
    pstamp(k, 2);
    acc_wait(k, 0);
    stream_prefetch(k, next_units);
    pstamp(k, 3);
    const float *pc = reinterpret_cast<const float *>(sm + O_PC) + 32 * k.wh;
    const uint32_t lane_addr = k.tm + ((uint32_t)(k.sub * 32) << 16) + 32 * k.wh;
    float z[32];
    {
        float c1[32], c2[32];
        tmem_ld32(lane_addr + T_MAIN, z);
        tmem_ld32(lane_addr + T_C1, c1);
        tmem_ld32(lane_addr + T_C2, c2);
#pragma unroll
        for (int e = 0; e < 32; ++e) z[e] = (z[e] + (c1[e] + c2[e])) + pc[e];
    }
    tc::tc_fence_before();
    // LayerNorm2 statistics: (mean, M2) over this thread's 32 columns, merged over the row's 16 pieces (Chan, fixed order)
    float m = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) m += z[e];
    m *= (1.f / 32.f);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const float d = z[e] - m;
        q = fmaf(d, d, q);
    }
    pstamp(k, 4);
    exchange_push(k, m, q, 0.f, 0.f, true);
    pstamp(k, 5);
    float mean = 0.f, m2 = 0.f;
#pragma unroll
    for (int s = 0; s < CL; ++s) mean += exchange_get(k, s).x;
    mean *= (1.f / CL);
#pragma unroll
    for (int s = 0; s < CL; ++s) {
        const float4 p = exchange_get(k, s);
        const float d = p.x - mean;
        m2 += fmaf(64.f * d, d, p.y);
    }
    rstd = 1.f / sqrtf(m2 * (1.f / H2) + LN_EPS);
    float p[L::nout];
#pragma unroll
    for (int o = 0; o < L::nout; ++o) p[o] = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        xh[e] = (z[e] - mean) * rstd;
        const float h = act_f(fmaf(xh[e], pc[64 + e], pc[128 + e]), slope);
#pragma unroll
        for (int o = 0; o < L::nout; ++o) p[o] = fmaf(h, pc[192 + 64 * o + e], p[o]);
    }
    exchange_push(k, p[0], L::nout > 1 ? p[1 % L::nout] : 0.f, L::nout > 2 ? p[2 % L::nout] : 0.f, L::nout > 3 ? p[3 % L::nout] : 0.f, false);
#pragma unroll
    for (int o = 0; o < L::nout; ++o) out[o] = 0.f;
#pragma unroll
    for (int s = 0; s < CL; ++s) {
        const float4 t = exchange_get(k, s);
        out[0] += t.x;
        if (L::nout > 1) { out[1 % L::nout] += t.y; out[2 % L::nout] += t.z; out[3 % L::nout] += t.w; }
    }
    const float *b3 = reinterpret_cast<const float *>(sm + O_PC) + 448;
#pragma unroll
    for (int o = 0; o < L::nout; ++o) out[o] += b3[o];
    pstamp(k, 6);
}

// ---- backward through head + LayerNorm2 of this thread's 32 columns --------------------------------------------------
// in: xh, rstd, dout[nout].  out: dz (grad of the pre-LayerNorm row), dy (grad at the LayerNorm output), h (activations)
template <class L>
__device__ __forceinline__ void head_ln2_backward(Ctl &k, const float (&xh)[32], float rstd, const float (&dout)[L::nout], float slope,
                                                  float (&dz)[32], float (&dy)[32], float (&h)[32]) {
    const float *pc = reinterpret_cast<const float *>(k.sm + O_PC) + 32 * k.wh;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) {
        const float pre = fmaf(xh[e], pc[64 + e], pc[128 + e]);
        h[e] = act_f(pre, slope);
        float g = 0.f;
#pragma unroll
        for (int o = 0; o < L::nout; ++o) g = fmaf(dout[o], pc[192 + 64 * o + e], g);
        dy[e] = g * act_g(pre, slope);
        dz[e] = dy[e] * pc[64 + e];
        s1 += dz[e];
        s2 = fmaf(dz[e], xh[e], s2);
    }
    exchange_push(k, s1, s2, 0.f, 0.f, false);
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int s = 0; s < CL; ++s) {
        const float4 p = exchange_get(k, s);
        t1 += p.x;
        t2 += p.y;
    }
    t1 *= (1.f / H2);
    t2 *= (1.f / H2);
#pragma unroll
    for (int e = 0; e < 32; ++e) dz[e] = rstd * (dz[e] - t1 - xh[e] * t2);
}

// column sums over the 64 rows of the block of up to 7 per-thread arrays (second-layer small parameters), staged through
// the (dead) H1 region.  val(a, e): value of array a at this thread's column e.  Result a, column c -> dst[a * dst_stride + c].
// Up to 8 per-row scalars (rv) are summed over the rows as well: sum j -> rdst[j] (skipped if rdst == nullptr).
constexpr int CS_LD = 68;
template <int NA, int NR, class F>
__device__ __forceinline__ void column_sums(Ctl &k, F val, float *dst, int dst_stride, const float (&rv)[NR], float *rdst) {
    float *st = reinterpret_cast<float *>(k.sm + O_R), *rs = st + NA * RB * CS_LD;
    if (k.act) {
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            float *row = st + (a * RB + k.row) * CS_LD + 32 * k.wh;
#pragma unroll
            for (int e = 0; e < 32; e += 4) st4(row + e, make_float4(val(a, e), val(a, e + 1), val(a, e + 2), val(a, e + 3)));
        }
        if (k.wh == 0) {
#pragma unroll
            for (int j = 0; j < NR; ++j) rs[j * RB + k.row] = rv[j];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NA * NSL + NR; i += TTHREADS) {
        if (i < NA * NSL) {
            const int a = i >> 6, c = i & 63;
            const float *col = st + a * RB * CS_LD + c;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;               // four independent chains (fixed order): the loop is latency-bound
#pragma unroll 4
            for (int r = 0; r < RB; r += 4) {
                s0 += col[r * CS_LD]; s1 += col[(r + 1) * CS_LD]; s2 += col[(r + 2) * CS_LD]; s3 += col[(r + 3) * CS_LD];
            }
            dst[a * dst_stride + c] = (s0 + s1) + (s2 + s3);
        } else if (rdst) {
            const int j = i - NA * NSL;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
            for (int r = 0; r < RB; r += 4) {
                s0 += rs[j * RB + r]; s1 += rs[j * RB + r + 1]; s2 += rs[j * RB + r + 2]; s3 += rs[j * RB + r + 3];
            }
            rdst[j] = (s0 + s1) + (s2 + s3);
        }
    }
    __syncthreads();
}

// ---- backward GEMMs of one pass ---------------------------------------------------------------------------------------
// dz: this thread's 32 columns of dz2.  Produces the LayerNorm1-backward result of the CTA's own 8 rows (one per warp) and,
// if WITH_DW, this row block's dW2 partial of the CTA's 64 features (from the H1^T image h1t in L2).
struct BwdOut { float dz1[8], dy1[8]; int b; };   // own row (warp w <-> row 8 crank + w), lane owns n = lane + 32 j
template <class L, bool WITH_DW>
__device__ __forceinline__ void pass_backward(Ctl &k, const float *tw, const float *imgt, float slope, const float (&dz)[32], const OwnRow &own,
                                              float *dh1p_unit, const float *h1t, float *gw2p_slot, BwdOut &o, const float *next_units) {
    char *sm = k.sm;
    jitter(k);
    const float *units = slice_of(k, imgt);                           // 8 sub-units: (mb, cc, hi | lo)
    stream_begin(k, units);
    if (WITH_DW && threadIdx.x == PRODUCER) {                         // H1^T image of the row block -> the (dead) H1 region
        fence_proxy_async_all();
        mbar_expect_tx(k.bars + B_H1T, 131072);
#pragma unroll
        for (int q = 0; q < 4; ++q) bulk_g2s(sm + O_R + q * 32768, h1t + (size_t)q * 8192, 32768, k.bars + B_H1T);
    }
    // (1) dz2 tile, K-major: block wh, row = batch row
    if (k.act) {
        char *dh = sm + O_D + k.wh * 8192 + (k.row >> 3) * 1024 + (k.row & 7) * 128;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
            float4 hi, lo;
            hi.x = tf32_hi(dz[4 * c4]); hi.y = tf32_hi(dz[4 * c4 + 1]); hi.z = tf32_hi(dz[4 * c4 + 2]); hi.w = tf32_hi(dz[4 * c4 + 3]);
            lo.x = dz[4 * c4] - hi.x; lo.y = dz[4 * c4 + 1] - hi.y; lo.z = dz[4 * c4 + 2] - hi.z; lo.w = dz[4 * c4 + 3] - hi.w;
            const int off = (c4 ^ (k.row & 7)) << 4;
            *reinterpret_cast<float4 *>(dh + off) = hi;
            *reinterpret_cast<float4 *>(dh + 16384 + off) = lo;
        }
    }
    pstamp(k, 0);
    publish_operands();
    stream_units(k, units, NSLOT, 8 - NSLOT);
    // (2) dH1^T[256][64] = W2T-image[256][64 features] . dz2[64 rows][64 features]^T : M = 128 x 2, N = 64, K = 64.
    //     issuer p: 0 = Ahi.Bhi, 1 = Alo.Bhi, 2 = Ahi.Blo; accumulator (mb, p); every issuer commits every sub-unit
    if (k.lane == 0 && k.warp < ISSUERS) {
        tc::tc_fence_after();
        const int p = k.warp;
        const uint32_t id = idesc(128, RB, 0, 0), d = smem_u32(sm + O_D) + (p == 2 ? 16384 : 0), ring = smem_u32(sm + O_RING);
        for (int mb = 0; mb < 2; ++mb)
            for (int cc = 0; cc < 2; ++cc) {
                const uint32_t acc = k.tm + 64 * (3 * mb + p);
                const uint64_t b = desc_k(d + cc * 8192);
                for (int part = 0; part < 2; ++part) {
                    const unsigned n = k.ring_n + (unsigned)(((mb * 2 + cc) * 2) + part);
                    ring_wait_full(k, n);                              // every issuer, also for sub-units it does not read: keeps its commits in step with the ring
                    if ((p == 1) == (part == 1)) {
                        const uint64_t a = desc_k(ring + (n % NSLOT) * SLOT_BYTES);
#pragma unroll
                        for (int s = 0; s < 4; ++s) mma_ss(acc, a + 2 * s, b + 2 * s, id, (cc == 0 && s == 0) ? 0u : 1u);
                    }
                    tc::mma_commit(k.bars + B_EMPTY + (n % NSLOT));
                }
            }
        tc::mma_commit(k.bars + B_ACC + 0);
    }
    __syncwarp();
    k.ring_n += 8u;
    pstamp(k, 1);
    acc_wait(k, 0);                                                   // dH1^T complete: the ring and the dz2 tile are idle
    if (!WITH_DW) stream_prefetch(k, next_units);
    pstamp(k, 2);
    // (3) dH1^T partial -> global [64 rows][256], coalesced along the 256 (lane = feature of H1)
    {
        const uint32_t lane_addr = k.tm + ((uint32_t)(k.sub * 32) << 16) + 32 * k.wh;
        float *dst = dh1p_unit + (size_t)k.crank * RB * H1;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            float v[32], c1[32], c2[32];
            tmem_ld32(lane_addr + 64 * (3 * mb), v);
            tmem_ld32(lane_addr + 64 * (3 * mb + 1), c1);
            tmem_ld32(lane_addr + 64 * (3 * mb + 2), c2);
            const int n = 128 * mb + 32 * k.sub + k.lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)(32 * k.wh + i) * H1 + n] = v[i] + (c1[i] + c2[i]);
        }
        tc::tc_fence_before();
    }
    pstamp(k, 3);
    if (WITH_DW) {
        // (4) dz2^T tile [64 features][64 rows] into the ring region (K = batch row), re-read from the K-major tile (hi + lo is exact)
        if (k.act) {
            const char *dh = sm + O_D + k.wh * 8192 + (k.row >> 3) * 1024 + (k.row & 7) * 128;
            char *e0 = sm + O_RING + (k.row >> 5) * 8192;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
                const int off = (c4 ^ (k.row & 7)) << 4;
                const float4 hi = *reinterpret_cast<const float4 *>(dh + off), lo = *reinterpret_cast<const float4 *>(dh + 16384 + off);
                const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t w = sw128_off(32 * k.wh + 4 * c4 + q, k.row & 31);
                    *reinterpret_cast<float *>(e0 + w) = hv[q];
                    *reinterpret_cast<float *>(e0 + 16384 + w) = lv[q];
                }
            }
        }
        publish_operands();                                           // also: every thread has read its dH1 accumulators (TMEM is reused)
        // (5) dW2^T[256][64] = H1^T[256][64 rows] . dz2^T[64 features][64 rows]^T : M = 128 x 2, N = 64, K = 64
        if (k.lane == 0 && k.warp < ISSUERS) {
            tc::mbar_wait(k.bars + B_H1T, k.h1t_n & 1u, 950);
            tc::tc_fence_after();
            const int p = k.warp;
            const uint32_t id = idesc(128, NSL, 0, 0), r = smem_u32(sm + O_R) + (p == 1 ? R_LO : 0), e = smem_u32(sm + O_RING) + (p == 2 ? 16384 : 0);
            for (int mb = 0; mb < 2; ++mb) {
                const uint32_t acc = k.tm + 64 * (3 * mb + p);
                for (int bb = 0; bb < 2; ++bb) {
                    const uint64_t a = desc_k(r + bb * 32768 + mb * 16384), b = desc_k(e + bb * 8192);
#pragma unroll
                    for (int s = 0; s < 4; ++s) mma_ss(acc, a + 2 * s, b + 2 * s, id, (bb == 0 && s == 0) ? 0u : 1u);
                }
            }
            tc::mma_commit(k.bars + B_ACC + 1);
        }
        __syncwarp();
        k.h1t_n += 1u;
    }
    pstamp(k, 4);
    jitter(k, true);
    cluster_sync();                                                   // every CTA's dH1 partial is visible (cluster-scope release / acquire)
    pstamp(k, 5);
    // (6) the CTA's own 8 rows: sum the 8 partials in a fixed order, LayerNorm1 backward (xhat1 / rstd1 kept from the forward)
    {
        const int b = 8 * k.crank + k.warp, lane = k.lane;
        o.b = b;
        float dh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dh[j] = 0.f;
#pragma unroll
        for (int src = 0; src < CL; ++src) {
            const float *p = dh1p_unit + ((size_t)src * RB + b) * H1;
#pragma unroll
            for (int j = 0; j < 8; ++j) dh[j] += __ldcg(p + lane + 32 * j);
        }
        ln1_backward<L>(tw, dh, own.xh1, own.rstd1, slope, lane, o.dz1, o.dy1);
    }
    pstamp(k, 6);
    if (WITH_DW) {
        // (7) dW2^T partial of this row block -> global [512][256] slot, rows 64 crank .. + 63
        acc_wait(k, 1);
        stream_prefetch(k, next_units);                               // the ring held the dz2^T tile until now
        const uint32_t lane_addr = k.tm + ((uint32_t)(k.sub * 32) << 16) + 32 * k.wh;
        float *dst = gw2p_slot + (size_t)NSL * k.crank * H1;
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
            float v[32], c1[32], c2[32];
            tmem_ld32(lane_addr + 64 * (3 * mb), v);
            tmem_ld32(lane_addr + 64 * (3 * mb + 1), c1);
            tmem_ld32(lane_addr + 64 * (3 * mb + 2), c2);
            const int kcol = 128 * mb + 32 * k.sub + k.lane;
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(size_t)(32 * k.wh + i) * H1 + kcol] = v[i] + (c1[i] + c2[i]);
        }
        tc::tc_fence_before();
    }
    __syncthreads();
    pstamp(k, 7);
}

// first-layer parameter gradients of the CTA's 8 rows (one per warp) -> partial slot (L layout, first-layer section)
template <class L>
__device__ __forceinline__ void ln1_param_partials(Ctl &k, const BwdOut &o, const OwnRow &own, int soff, const float *acts, float *part) {
    float *t_dz = reinterpret_cast<float *>(k.sm + O_D), *t_dg = t_dz + 8 * H1, *t_dbe = t_dg + 8 * H1, *t_x = t_dbe + 8 * H1;   // [8][256] x 3 + [8][32]
    const float *xs = reinterpret_cast<const float *>(k.sm + O_X);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int n = k.lane + 32 * j;
        t_dz[k.warp * H1 + n] = o.dz1[j];
        t_dg[k.warp * H1 + n] = o.dy1[j] * own.xh1[j];
        t_dbe[k.warp * H1 + n] = o.dy1[j];
    }
    if (k.lane < L::k0) t_x[k.warp * 32 + k.lane] = k.lane < S ? xs[o.b * 32 + soff + k.lane] : (acts ? acts[o.b * 4 + (k.lane - S)] : xs[o.b * 32 + k.lane]);
    __syncthreads();
    {
        const int n = threadIdx.x;   // 256 threads = 256 columns
        float db1 = 0.f, dg1 = 0.f, dbe1 = 0.f, dw[L::k0];
#pragma unroll
        for (int kk = 0; kk < L::k0; ++kk) dw[kk] = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const float dzv = t_dz[w * H1 + n];
            db1 += dzv;
            dg1 += t_dg[w * H1 + n];
            dbe1 += t_dbe[w * H1 + n];
#pragma unroll
            for (int kk = 0; kk < L::k0; ++kk) dw[kk] = fmaf(dzv, t_x[w * 32 + kk], dw[kk]);
        }
        part[L::B1 + n] = db1;
        part[L::G1 + n] = dg1;
        part[L::BE1 + n] = dbe1;
#pragma unroll
        for (int kk = 0; kk < L::k0; ++kk) part[L::W1T + kk * H1 + n] = dw[kk];
    }
    __syncthreads();
}

// ---- gather ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gather_rows(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step, int r, bool bc_rows) {
    float *xs = reinterpret_cast<float *>(k.sm + O_X);
    const int32_t *ion = c.idx_online + (size_t)step * c.B, *iex = c.idx_expert + (size_t)step * c.B, *ibc = c.idx_bc + (size_t)step * c.B;
    const float *src[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {                                     // all index loads, then all row loads: independent round trips in flight
        int b = RB * r + k.warp + 8 * i;
        if (b >= c.B) b = c.B - 1;                                   // padded rows repeat the last row; their contributions are masked
        if (bc_rows) src[i] = c.bctab + (size_t)__ldg(ibc + b) * BCW;
        else src[i] = b < st.n_online ? c.online + (size_t)__ldg(ion + b) * ROW : c.expert + (size_t)__ldg(iex + (b - st.n_online)) * ROW;
    }
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = (!bc_rows || k.lane < BCW) ? __ldg(src[i] + k.lane) : 0.f;
    __syncthreads();                                                  // previous readers of the row buffer are done
#pragma unroll
    for (int i = 0; i < 8; ++i) xs[(k.warp + 8 * i) * 32 + k.lane] = w[i];
    __syncthreads();
}

// ---- Adam over one network (all CTAs of the grid), fixed-order sums of the partials, operand images ------------------
//   w2parts [tower][nparts][H2*H1]; part: slots of L::SMALL floats, slot s of tower t at ((s * ntowers) + t) * L::SMALL;
//   first-layer entries are summed over slots 0 .. n1-1, second-layer small entries over slots 0, 8, 16, .. (nq of them).
//   W2 is processed in tiles of [32 features][16 k]: reads / arena writes / forward image coalesced along k, the
//   input-gradient image (features contiguous) through a shared-memory transpose.
template <class L, int NPARTS, bool BRANCHES, int UB>
__device__ __forceinline__ void adam_tcc(Ctl &k, int rank, int size, int ntowers, float *w, float *gw, const float *w2parts, float *m, float *v,
                                         const float *part, int n1, int nq, float neg_step, float bc2_sqrt, float *img, float *imgt,
                                         float *target, float *timg, bool polyak, float tau, float omt, float sc0 = 1.f, float sc1 = 1.f) {
    // BRANCHES: the partial slots are two halves (RL branch | cloning branch) whose sums are scaled by sc0 / sc1
    // (a) everything but W2: runs between the loads and the arithmetic of the first W2 tile round (both are bound by L2
    //     round trips, so their latencies overlap instead of adding up)
    const int nsmall = ntowers * L::SMALL;
    bool small_done = false;
    auto small_params = [&]() {
    small_done = true;
    for (int i = rank * TTHREADS + threadIdx.x; i < nsmall; i += size * TTHREADS) {
        const int tower = i / L::SMALL, si = i - tower * L::SMALL;
        const int off = si < L::W2 ? si : si + H2 * H1, e = tower * L::SIZE + off;
        const bool first = off < L::W2;
        const int ns = first ? n1 : nq, stride = first ? 1 : CL;
        const float *p0 = part + (size_t)tower * L::SMALL + si;
        const size_t ps = (size_t)stride * ntowers * L::SMALL;
        float g = 0.f;
        if (BRANCHES) {
            float g1 = 0.f;
#pragma unroll 8
            for (int s = 0; s < ns / 2; ++s) {
                g += __ldcg(p0 + s * ps);
                g1 += __ldcg(p0 + (s + ns / 2) * ps);
            }
            g = fmaf(sc0, g, sc1 * g1);
        } else {
#pragma unroll 8
            for (int s = 0; s < ns; ++s) g += __ldcg(p0 + s * ps);
        }
        gw[e] = g;
        float wv = w[e], mv = m[e], vv = v[e];
        adam_elem(g, wv, mv, vv, neg_step, bc2_sqrt);
        w[e] = wv; m[e] = mv; v[e] = vv;
        if (polyak) target[e] = polyak_elem(target[e], wv, tau, omt);
    }
    };
    // (b) W2 tiles, up to 3 tiles' loads in flight per CTA (the loop is bound by L2 round trips, not by arithmetic)
    float *ts = reinterpret_cast<float *>(k.sm + O_R);                 // [32 features][17]
    const int tiles = ntowers * 16 * 16, tk = threadIdx.x & 15, tn = threadIdx.x >> 4;
    for (int t0 = rank; t0 < tiles; t0 += UB * size) {
        float g[UB][2], wv[UB][2], mv[UB][2], vv[UB][2], tv[UB][2];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int t = t0 + u * size;
            if (t < tiles) {
                const int tower = t >> 8, nb = (t >> 4) & 15, kb = t & 15;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int n = 32 * nb + tn + 16 * it, kk = 16 * kb + tk, i = n * H1 + kk, e = tower * L::SIZE + L::W2 + i;
                    const float *p0 = w2parts + (size_t)tower * NPARTS * H2 * H1 + i;
                    float s = __ldcg(p0);
                    if (BRANCHES) {
                        float s1 = __ldcg(p0 + (size_t)(NPARTS / 2) * H2 * H1);
#pragma unroll
                        for (int q = 1; q < NPARTS / 2; ++q) {
                            s += __ldcg(p0 + (size_t)q * H2 * H1);
                            s1 += __ldcg(p0 + (size_t)(q + NPARTS / 2) * H2 * H1);
                        }
                        s = fmaf(sc0, s, sc1 * s1);
                    } else {
#pragma unroll
                        for (int q = 1; q < NPARTS; ++q) s += __ldcg(p0 + (size_t)q * H2 * H1);
                    }
                    g[u][it] = s;
                    wv[u][it] = w[e]; mv[u][it] = m[e]; vv[u][it] = v[e];
                    tv[u][it] = polyak ? target[e] : 0.f;
                }
            }
        }
        if (!small_done) small_params();
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int t = t0 + u * size;
            if (t < tiles) {
                const int tower = t >> 8, nb = (t >> 4) & 15, kb = t & 15;
#pragma unroll
                for (int it = 0; it < 2; ++it) {
                    const int n = 32 * nb + tn + 16 * it, kk = 16 * kb + tk, i = n * H1 + kk, e = tower * L::SIZE + L::W2 + i;
                    gw[e] = g[u][it];
                    adam_elem(g[u][it], wv[u][it], mv[u][it], vv[u][it], neg_step, bc2_sqrt);
                    w[e] = wv[u][it]; m[e] = mv[u][it]; v[e] = vv[u][it];
                    const float hi = tf32_hi(wv[u][it]);
                    float *f = img + (size_t)tower * IMG_FWD + img_fwd_index(n, kk);
                    f[0] = hi; f[2048] = wv[u][it] - hi;
                    if (polyak) {
                        const float tnew = polyak_elem(tv[u][it], wv[u][it], tau, omt);
                        target[e] = tnew;
                        const float th = tf32_hi(tnew);
                        float *f2 = timg + (size_t)tower * IMG_FWD + img_fwd_index(n, kk);
                        f2[0] = th; f2[2048] = tnew - th;
                    }
                }
                if (imgt) {
                    __syncthreads();
                    ts[tn * 17 + tk] = wv[u][0];
                    ts[(tn + 16) * 17 + tk] = wv[u][1];
                    __syncthreads();
                    const int n2 = threadIdx.x & 31, k2 = threadIdx.x >> 5;   // 32 features x 8 k per sweep
#pragma unroll
                    for (int it = 0; it < 2; ++it) {
                        const int kk = 16 * kb + k2 + 8 * it, n = 32 * nb + n2;
                        const float x = ts[n2 * 17 + k2 + 8 * it], hi = tf32_hi(x);
                        float *tq = imgt + (size_t)tower * IMG_T + img_t_index(n, kk);
                        tq[0] = hi; tq[4096] = x - hi;
                    }
                }
            }
        }
    }
    if (!small_done) small_params();
}
// Polyak of a network whose online copy did not change in this phase (the critic, in phase D)
template <class L>
__device__ __forceinline__ void polyak_tcc(int rank, int size, int ntowers, const float *w, float *target, float *timg, float tau, float omt) {
    const int total = ntowers * L::SIZE;
    for (int e = rank * TTHREADS + threadIdx.x; e < total; e += size * TTHREADS) {
        const int tower = e / L::SIZE, off = e - tower * L::SIZE;
        const float tv = polyak_elem(target[e], w[e], tau, omt);
        target[e] = tv;
        if (off >= L::W2 && off < L::B2) {
            const int i = off - L::W2, n = i / H1, kk = i - n * H1;
            const float th = tf32_hi(tv);
            float *f2 = timg + (size_t)tower * IMG_FWD + img_fwd_index(n, kk);
            f2[0] = th; f2[2048] = tv - th;
        }
    }
}

// rebuilds the operand images of one tower from its arena (after hirl_params_import)
__global__ void tcc_images_kernel(const float *w2, float *img, float *imgt) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < H2 * H1; i += gridDim.x * blockDim.x) {
        const int n = i / H1, kk = i - n * H1;
        const float wv = w2[i], hi = tf32_hi(wv);
        float *f = img + img_fwd_index(n, kk);
        f[0] = hi; f[2048] = wv - hi;
        if (imgt) {
            float *t = imgt + img_t_index(n, kk);
            t[0] = hi; t[4096] = wv - hi;
        }
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TTHREADS, 1) hirl_tcc_kernel(const __grid_constant__ TccLaunch Lc) {
    extern __shared__ float4 smem4[];
    __shared__ AgentCtx ctx;
    __shared__ hirl_step_t st;
    __shared__ uint64_t bars[N_BARS];
    __shared__ uint32_t tmem_slot;
    {
        const int *src = reinterpret_cast<const int *>(&Lc.agent);
        int *dst = reinterpret_cast<int *>(&ctx);
        for (int i = threadIdx.x; i < (int)(sizeof(AgentCtx) / sizeof(int)); i += TTHREADS) dst[i] = src[i];
    }
    Ctl k;
    k.pad = (1024u - (smem_u32(smem4) & 1023u)) & 1023u;
    k.sm = reinterpret_cast<char *>(smem4) + k.pad;           // plain pointer arithmetic on the shared symbol keeps the address space
    k.bars = bars;
    k.crank = (int)cluster_ctarank(); k.cid = (int)cluster_id_x(); k.ncl = Lc.ncl;
    k.warp = threadIdx.x >> 5; k.lane = threadIdx.x & 31; k.sub = k.warp & 3; k.wh = k.warp >> 2;
    k.act = k.lane < 16; k.row = 16 * k.sub + (k.lane & 15);
    k.ring_n = 0u; k.acc_n[0] = k.acc_n[1] = 0u; k.w1_n = k.h1t_n = k.h1_n = 0u; k.xr = 0u;
    k.flags = Lc.bar_flags;
    k.gen = ld_relaxed_u32(k.flags + k.cid * FLAG_STRIDE);
    k.pr = k.pr2 = nullptr; k.pb = 0; k.pre = nullptr;
    k.jit = Lc.jitter; k.jit_n = 0u; k.jit_ns = Lc.jitter_ns ? Lc.jitter_ns : 4000u;
    if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot);
    if (threadIdx.x == 32) {
        for (int i = 0; i < N_BARS; ++i) tc::mbar_init(bars + i, (i >= B_EMPTY && i < B_ACC + 2) ? ISSUERS : 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    k.tm = tmem_slot;
    const AgentCtx &c = ctx;
    k.pflags = c.tcc.pflags;
    k.pgen = ld_relaxed_u32(c.tcc.gen_ctr);
    k.h1img = c.tcc.H1I + (size_t)k.cid * H1T_FLOATS;
    cluster_sync();                                                    // peers' shared memory and barriers are live before anybody pushes
    const float slope = slope_of(c);
    const int B = c.B, nrb = (B + RB - 1) / RB;
    float *acts = reinterpret_cast<float *>(k.sm + O_ACT), *gpi = reinterpret_cast<float *>(k.sm + O_GPI);
    const int grank = k.cid * CL + k.crank, gsize = k.ncl * CL;
    const float tau = c.tau, omt = (float)(1.0 - (double)c.tau);

    for (int step = 0; step < Lc.k_steps; ++step) {
        __syncthreads();
        if (threadIdx.x < (int)(sizeof(hirl_step_t) / sizeof(int)))
            reinterpret_cast<int *>(&st)[threadIdx.x] = reinterpret_cast<const int *>(c.steps + step)[threadIdx.x];
        __syncthreads();
        const bool do_actor = st.do_actor != 0, soft = st.bc_weight_now == 100.f, do_polyak = st.do_polyak != 0;
        const unsigned fgen = k.pgen + (unsigned)step + 1u;
        k.pr = (Lc.prof && step < PROF_STEPS && k.cid == 0 && k.crank == 0) ? Lc.prof + (size_t)step * PROF_WORDS : nullptr;
        k.pr2 = (Lc.prof && step < PROF_STEPS && k.cid == nrb % k.ncl && k.crank == 0) ? Lc.prof + (size_t)step * PROF_WORDS : nullptr;
        stamp(k, 0);
        stamp2(k, 18);
        if (k.pr && threadIdx.x == 0) k.pr[19] = k.ncl;
        // =============================== phase A: critic forward / backward ===============================
        // units 0 .. nrb - 1: target actor of row block r -> a' published through L2;
        // units nrb .. 3 nrb - 1: critic tower (r, t): forward on (s, a) while the target actor runs, then the target tower on
        // (s', a'), the TD target with the partner tower's target Q, loss and the backward pass.
        // 3 nrb <= 12 clusters carry the phase (B200 co-schedules 15 clusters of 8 CTAs).
        //   on actor calls, units 3 nrb .. 4 nrb - 1: ACTOR forward unit r: pi = actor(s) (HIRL.py:295) depends on nothing phase A or B
        //   changes, so it runs here on clusters that would idle (its state is parked in L2 for phase C's role A)
        const int nunits_a = 3 * nrb + (do_actor ? nrb : 0);
#pragma unroll 1
        for (int ua = k.cid; ua < nunits_a; ua += k.ncl) {
            float xh[32], rstd;
            OwnRow own;
            if (ua >= 3 * nrb) {
                const int r = ua - 3 * nrb;
                const int bg = RB * r + k.row;
                gather_rows(k, c, st, step, r, false);
                float out[A];
                pass_forward<ActorL>(k, c.actor, c.tcc.img_actor, 0, nullptr, slope, c.tcc.H1TI + (size_t)(12 + r) * H1T_FLOATS, xh, rstd, out, own, nullptr);
                if (k.crank == 0 && k.act && k.wh == 0) {
                    const float4 pi = make_float4(tanhf(out[0]), tanhf(out[1]), tanhf(out[2]), tanhf(out[3]));
                    st4(c.tcc.PIX + (r * RB + k.row) * 4, pi);
                    c.tcc.RSA[r * RB + k.row] = rstd;
                    if (bg < B) st4(c.sc.PI + (size_t)bg * A, pi);
                }
                if (k.act) {
                    float *xs2 = c.tcc.XH2S + ((size_t)r * RB + k.row) * H2 + NSL * k.crank + 32 * k.wh;
#pragma unroll
                    for (int e = 0; e < 32; e += 4) st4(xs2 + e, make_float4(xh[e], xh[e + 1], xh[e + 2], xh[e + 3]));
                }
                {
                    const int b = 8 * k.crank + k.warp;                 // the row this warp owns
                    float *op = c.tcc.OWNP + ((size_t)r * RB + b) * (H1 + 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) op[k.lane + 32 * j] = own.xh1[j];
                    if (k.lane == 0) op[H1] = own.rstd1;
                }
                continue;
            }
            if (ua < nrb) {
                const int r = ua;
                gather_rows(k, c, st, step, r, false);
                stamp(k, 1);
                // target actor on next_state (HIRL.py:264) -> a' (HIRL.py:265-267)
                float out[A];
                k.pb = 20;
                pass_forward<ActorL>(k, c.tactor, c.tcc.img_tactor, SA, nullptr, slope, nullptr, xh, rstd, out, own, nullptr);
                k.pb = 0;
                if (k.crank == 0 && k.act && k.wh == 0) {
                    float ap[A];
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float nz = fminf(fmaxf(st.noise[j], -c.noise_clamp), c.noise_clamp);
                        ap[j] = fminf(fmaxf(tanhf(out[j]) + nz, -1.f), 1.f);
                    }
                    st4(c.tcc.APX + (r * RB + k.row) * 4, make_float4(ap[0], ap[1], ap[2], ap[3]));
                }
                flag_publish(k, PF_AP + r, fgen);
                stamp(k, 2);
                continue;
            }
            const int unit = ua - nrb, r = unit >> 1, t = unit & 1;
            const int bg = RB * r + k.row;                             // global batch row of this thread
            const bool valid = bg < B;
            gather_rows(k, c, st, step, r, false);
            stamp2(k, 17);
            // critic tower t on (state, action) (HIRL.py:277)
            float q1[1];
            float *h1t = c.tcc.H1TI + (size_t)unit * H1T_FLOATS;
            k.pb = 28;
            pass_forward<CriticL>(k, c.critic + t * CriticL::SIZE, c.tcc.img_critic + (size_t)t * IMG_FWD, 0, nullptr, slope, h1t, xh, rstd, q1, own,
                                  slice_of(k, c.tcc.img_tcritic + (size_t)t * IMG_FWD));
            k.pb = 0;
            stamp2(k, 4);
            float *xpark = c.tcc.XHP + ((size_t)unit * RB + k.row) * H2 + NSL * k.crank + 32 * k.wh;   // the pass' state waits in L2 while the target tower uses the registers
            if (k.act) {
#pragma unroll
                for (int e = 0; e < 32; e += 4) st4(xpark + e, make_float4(xh[e], xh[e + 1], xh[e + 2], xh[e + 3]));
            }
            const float rstd_c = rstd;
            // target critic tower t on (next_state, a') (HIRL.py:270)
            flag_wait(k, PF_AP + r, fgen);
            if (k.act && k.wh == 0) *reinterpret_cast<float4 *>(acts + k.row * 4) = __ldcg(reinterpret_cast<const float4 *>(c.tcc.APX + (r * RB + k.row) * 4));
            stamp2(k, 38);
            float tq_own;
            {
                float out[1];
                OwnRow own_t;
                pass_forward<CriticL>(k, c.tcritic + t * CriticL::SIZE, c.tcc.img_tcritic + (size_t)t * IMG_FWD, SA, acts, slope, nullptr, xh, rstd, out, own_t,
                                      slice_of(k, c.tcc.imgt_critic + (size_t)t * IMG_T));
                tq_own = out[0];
                if (k.crank == 0 && k.act && k.wh == 0) c.tcc.TQX[(r * 2 + t) * RB + k.row] = tq_own;
                flag_publish(k, PF_TQ + r * 2 + t, fgen);
            }
            stamp2(k, 39);
            load_pc<CriticL>(k, c.critic + t * CriticL::SIZE);         // the parameter cache held the target tower's (free since flag_publish's barrier);
            if (k.act) {                                               // these loads and the parked state travel while we wait for the partner
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    const float4 v4 = __ldcg(reinterpret_cast<const float4 *>(xpark + e));
                    xh[e] = v4.x; xh[e + 1] = v4.y; xh[e + 2] = v4.z; xh[e + 3] = v4.w;
                }
            }
            rstd = rstd_c;
            flag_wait(k, PF_TQ + r * 2 + (1 - t), fgen);               // (its barrier also publishes the parameter cache)
            stamp2(k, 36);
            const float tq_other = __ldcg(c.tcc.TQX + (r * 2 + (1 - t)) * RB + k.row);
            const float *xs = reinterpret_cast<const float *>(k.sm + O_X);
            const float rew = xs[k.row * 32 + 30], dn = xs[k.row * 32 + 31];
            const float tq = t == 0 ? fminf(tq_own, tq_other) : fminf(tq_other, tq_own);
            const float y = __fadd_rn(rew, __fmul_rn(__fmul_rn(c.gamma, tq), __fsub_rn(1.f, dn)));          // HIRL.py:274
            const float delta = q1[0] - y;
            if (k.crank == 0 && k.act && k.wh == 0 && valid) {
                c.sc.Q[(size_t)t * B + bg] = q1[0];
                c.sc.TQ[(size_t)t * B + bg] = tq_own;
                if (t == 0) c.sc.Y[bg] = y;
            }
            float dout[1] = {valid ? (2.f / (float)B) * delta : 0.f};                                          // mse_loss backward, mean reduction
            float dz[32], dy[32], h[32];
            head_ln2_backward<CriticL>(k, xh, rstd, dout, slope, dz, dy, h);
            // second-layer small parameters of this CTA's 64 features: db2, dgamma2, dbeta2, dW3 (+ db3, loss by CTA 0)
            float *part = c.tcc.PARTC + ((size_t)r * CL * 2 + t) * CriticL::SMALL;                             // slot 8 r (+ tower)
            {
                float *p2 = part + CriticL::small_index(CriticL::B2) + NSL * k.crank;
                const float dq = dout[0];
                const float rv[2] = {dq, valid ? delta * delta : 0.f};
                column_sums<4, 2>(k, [&](int a, int e) { return a == 0 ? dz[e] : a == 1 ? dy[e] * xh[e] : a == 2 ? dy[e] : dq * h[e]; }, p2, H2, rv,
                                  k.crank == 0 ? c.tcc.RSUM + 8 * unit : nullptr);
                if (k.crank == 0 && threadIdx.x == 0) {
                    part[CriticL::small_index(CriticL::B3)] = c.tcc.RSUM[8 * unit];
                    c.tcc.LPT[unit] = c.tcc.RSUM[8 * unit + 1];
                }
            }
            stamp2(k, 37);
            BwdOut bo;
            k.pb = 40;
            pass_backward<CriticL, true>(k, c.critic + t * CriticL::SIZE, c.tcc.imgt_critic + (size_t)t * IMG_T, slope, dz, own,
                                         c.tcc.DH1P + (size_t)unit * CL * RB * H1, h1t, c.tcc.GW2PC + ((size_t)t * MAX_RB + r) * H2 * H1, bo, nullptr);
            k.pb = 0;
            ln1_param_partials<CriticL>(k, bo, own, 0, nullptr, c.tcc.PARTC + ((size_t)(r * CL + k.crank) * 2 + t) * CriticL::SMALL);
            stamp2(k, 5);
        }
        grid_sync(k);
        stamp(k, 6);
        // =============================== phase B: critic Adam ===============================
        adam_tcc<CriticL, MAX_RB, false, 5>(k, grank, gsize, 2, c.critic, c.g_critic, c.tcc.GW2PC, c.m_critic, c.v_critic, c.tcc.PARTC, CL * nrb, nrb,
                          st.critic_neg_step, st.critic_bc2_sqrt, c.tcc.img_critic, c.tcc.imgt_critic, nullptr, nullptr, false, 0.f, 0.f);
        if (grank == gsize - 1 && threadIdx.x == 0) {
            float s0 = 0.f, s1 = 0.f;
            for (int r = 0; r < nrb; ++r) { s0 += __ldcg(c.tcc.LPT + 2 * r); s1 += __ldcg(c.tcc.LPT + 2 * r + 1); }
            c.carry[0] = __fadd_rn(s0 / (float)B, s1 / (float)B);                                              // HIRL.py:284
        }
        stamp(k, 7);
        grid_sync(k);
        stamp(k, 8);
        if (do_actor) {
            // =============================== phase C: actor forward / backward ===============================
            // The Q-filter weight w (HIRL.py:292-308) needs every row block's count, but both actor branches are LINEAR in their
            // loss scale: the RL branch is back-propagated with scale 1 and the cloning branch with 2 / (A B), and phase D
            // multiplies the summed partial gradients of a branch by (1 - w) resp. w * lambda -- nobody waits for the count.
            const int role = k.cid % 3, r = k.cid / 3;
            if (k.cid < 3 * nrb && (role != 1 || soft)) {
                const int bg = RB * r + k.row;
                const bool valid = bg < B;
                float xh[32], rstd;
                OwnRow own, own_q;
                float *h1t = c.tcc.H1TI + (size_t)(role == 0 ? 12 + r : k.cid) * H1T_FLOATS;   // role A: written by phase A's actor unit
                float *rsum = c.tcc.RSUM + 8 * k.cid;
                if (role == 0) {
                    // ---- role A: rl_Q = Q1(s, pi) with the updated critic (HIRL.py:296-297); pi = actor(s) was evaluated in phase A
                    gather_rows(k, c, st, step, r, false);
                    float pi[A];
                    {
                        const float4 p4 = __ldcg(reinterpret_cast<const float4 *>(c.tcc.PIX + (r * RB + k.row) * 4));
                        pi[0] = p4.x; pi[1] = p4.y; pi[2] = p4.z; pi[3] = p4.w;
                        if (k.act && k.wh == 0) *reinterpret_cast<float4 *>(acts + k.row * 4) = p4;
                    }
                    const float rstd_a = __ldcg(c.tcc.RSA + r * RB + k.row);
                    stamp(k, 10);
                    float q[1];
                    pass_forward<CriticL>(k, c.critic, c.tcc.img_critic, 0, acts, slope, nullptr, xh, rstd, q, own_q, slice_of(k, c.tcc.imgt_critic));
                    stamp(k, 11);
                    if (k.crank == 0 && k.act && k.wh == 0) {
                        c.tcc.RLQX[r * RB + k.row] = q[0];
                        if (valid) c.sc.RLQ[bg] = q[0];
                    }
                    if (soft) flag_publish(k, PF_RLQ + r, fgen);        // role S compares its soft Q with it (HIRL.py:303)
                    // backward of rl_loss = -mean(Q1) to the action input, unscaled: (1 - w) is applied to the branch's summed gradient in phase D
                    float dout[1] = {valid ? -1.f / (float)B : 0.f};
                    float dz[32], dy[32], h[32];
                    head_ln2_backward<CriticL>(k, xh, rstd, dout, slope, dz, dy, h);
                    {
                        const float rv[2] = {0.f, valid ? q[0] : 0.f};                                           // HIRL.py:297
                        column_sums<0, 2>(k, [&](int, int) { return 0.f; }, nullptr, 0, rv, k.crank == 0 ? rsum : nullptr);
                        if (k.crank == 0 && threadIdx.x == 0) c.tcc.CNT[2 * r + 1] = rsum[1];
                    }
                    BwdOut bo;
                    pass_backward<CriticL, false>(k, c.critic, c.tcc.imgt_critic, slope, dz, own_q, c.tcc.DH1P + (size_t)k.cid * CL * RB * H1, nullptr, nullptr, bo,
                                                  slice_of(k, c.tcc.imgt_actor));
                    {
                        float g[A];
#pragma unroll
                        for (int j2 = 0; j2 < A; ++j2) {
                            float p = 0.f;
#pragma unroll
                            for (int j = 0; j < 8; ++j) p = fmaf(bo.dz1[j], c.critic[CriticL::W1T + (S + j2) * H1 + k.lane + 32 * j], p);
                            g[j2] = warp_sum(p);
                        }
                        if (k.lane < CL) st_dsmem_f32x4(map_to_cta(smem_u32(gpi + bo.b * 4), k.lane), g[0], g[1], g[2], g[3]);
                        if (k.crank == 0 && k.lane == 0 && RB * r + bo.b < B) {
#pragma unroll
                            for (int j = 0; j < A; ++j) c.sc.GPI[(size_t)(RB * r + bo.b) * A + j] = g[j];
                        }
                        cluster_sync();
                    }
                    stamp(k, 12);
                    // actor backward, RL branch (state of the actor pass: parked by phase A's actor unit)
                    load_pc<ActorL>(k, c.actor);
                    {
                        const float *op = c.tcc.OWNP + ((size_t)r * RB + 8 * k.crank + k.warp) * (H1 + 32);
#pragma unroll
                        for (int j = 0; j < 8; ++j) own.xh1[j] = __ldcg(op + k.lane + 32 * j);
                        own.rstd1 = __ldcg(op + H1);
                    }
                    if (k.act) {
                        const float *xs2 = c.tcc.XH2S + ((size_t)r * RB + k.row) * H2 + NSL * k.crank + 32 * k.wh;
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float4 v4 = __ldcg(reinterpret_cast<const float4 *>(xs2 + e));
                            xh[e] = v4.x; xh[e + 1] = v4.y; xh[e + 2] = v4.z; xh[e + 3] = v4.w;
                        }
                    }
                    __syncthreads();
                    float dpre[A];
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float dpi = valid ? gpi[k.row * 4 + j] : 0.f;
                        dpre[j] = dpi * (1.f - pi[j] * pi[j]);
                    }
                    head_ln2_backward<ActorL>(k, xh, rstd_a, dpre, slope, dz, dy, h);
                    float *part = c.tcc.PARTA + (size_t)(r * CL) * ActorL::SMALL;                               // branch 0, slot 8 r
                    {
                        float *p2 = part + ActorL::small_index(ActorL::B2) + NSL * k.crank;
                        column_sums<7, 4>(k, [&](int a, int e) { return a == 0 ? dz[e] : a == 1 ? dy[e] * xh[e] : a == 2 ? dy[e] : dpre[a - 3] * h[e]; }, p2, H2,
                                          dpre, k.crank == 0 ? part + ActorL::small_index(ActorL::B3) : nullptr);
                    }
                    pass_backward<ActorL, true>(k, c.actor, c.tcc.imgt_actor, slope, dz, own, c.tcc.DH1P + (size_t)k.cid * CL * RB * H1, h1t,
                                                c.tcc.GW2PA + (size_t)r * H2 * H1, bo, nullptr);
                    ln1_param_partials<ActorL>(k, bo, own, 0, nullptr, c.tcc.PARTA + (size_t)(r * CL + k.crank) * ActorL::SMALL);
                    stamp(k, 14);
                } else if (role == 1) {
                    // ---- role S: a_bc = bc_actor(s), soft_Q = Q1(s, a_bc) (HIRL.py:300-302)
                    gather_rows(k, c, st, step, r, false);
                    float out[A];
                    pass_forward<ActorL>(k, c.bc, c.tcc.img_bc, 0, nullptr, slope, nullptr, xh, rstd, out, own, slice_of(k, c.tcc.img_critic));
                    if (k.act && k.wh == 0) {
#pragma unroll
                        for (int j = 0; j < A; ++j) {
                            const float a = tanhf(out[j]);
                            acts[k.row * 4 + j] = a;
                            if (k.crank == 0 && valid) c.sc.ABC[(size_t)bg * A + j] = a;
                        }
                    }
                    float q[1];
                    pass_forward<CriticL>(k, c.critic, c.tcc.img_critic, 0, acts, slope, nullptr, xh, rstd, q, own_q, nullptr);
                    // the count of soft_Q > rl_Q over the row block (HIRL.py:303); rl_Q comes from role A's unit
                    flag_wait(k, PF_RLQ + r, fgen);
                    const float rlq = __ldcg(c.tcc.RLQX + r * RB + k.row);
                    if (k.crank == 0 && k.act && k.wh == 0 && valid) c.sc.SOFTQ[bg] = q[0];
                    const float rv[2] = {(valid && q[0] > rlq) ? 1.f : 0.f, 0.f};
                    column_sums<0, 2>(k, [&](int, int) { return 0.f; }, nullptr, 0, rv, k.crank == 0 ? rsum : nullptr);
                    if (k.crank == 0 && threadIdx.x == 0) c.tcc.CNT[2 * r] = rsum[0];
                } else {
                    // ---- role AB: pi_bc = actor(s_bc), cloning loss (HIRL.py:310-319), actor backward of the BC branch
                    gather_rows(k, c, st, step, r, true);
                    float out[A], pi[A], diff[A];
                    pass_forward<ActorL>(k, c.actor, c.tcc.img_actor, 0, nullptr, slope, h1t, xh, rstd, out, own, slice_of(k, c.tcc.imgt_actor));
                    const float *xs = reinterpret_cast<const float *>(k.sm + O_X);
                    float sq = 0.f;
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        pi[j] = tanhf(out[j]);
                        diff[j] = pi[j] - xs[k.row * 32 + S + j];
                        sq = fmaf(diff[j], diff[j], sq);
                    }
                    const float fire = diff[3] * diff[3];
                    if (k.crank == 0 && k.act && k.wh == 0 && valid) {
#pragma unroll
                        for (int j = 0; j < A; ++j) c.sc.PIB[(size_t)bg * A + j] = pi[j];
                    }
                    {
                        const float rv[2] = {valid ? sq : 0.f, valid ? fire : 0.f};
                        column_sums<0, 2>(k, [&](int, int) { return 0.f; }, nullptr, 0, rv, k.crank == 0 ? c.tcc.LPT + 8 + 2 * r : nullptr);
                    }
                    float dpre[A];
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float dpi = valid ? (2.f / (float)(A * B)) * diff[j] : 0.f;
                        dpre[j] = dpi * (1.f - pi[j] * pi[j]);
                    }
                    float dz[32], dy[32], h[32];
                    head_ln2_backward<ActorL>(k, xh, rstd, dpre, slope, dz, dy, h);
                    float *part = c.tcc.PARTA + (size_t)((MAX_RB + r) * CL) * ActorL::SMALL;                    // branch 1, slot 8 (4 + r)
                    {
                        float *p2 = part + ActorL::small_index(ActorL::B2) + NSL * k.crank;
                        column_sums<7, 4>(k, [&](int a, int e) { return a == 0 ? dz[e] : a == 1 ? dy[e] * xh[e] : a == 2 ? dy[e] : dpre[a - 3] * h[e]; }, p2, H2,
                                          dpre, k.crank == 0 ? part + ActorL::small_index(ActorL::B3) : nullptr);
                    }
                    BwdOut bo;
                    cluster_sync();                                     // every CTA's rows of the H1^T image (written during this unit's forward MMAs) are visible
                    pass_backward<ActorL, true>(k, c.actor, c.tcc.imgt_actor, slope, dz, own, c.tcc.DH1P + (size_t)k.cid * CL * RB * H1, h1t,
                                                c.tcc.GW2PA + (size_t)(MAX_RB + r) * H2 * H1, bo, nullptr);
                    ln1_param_partials<ActorL>(k, bo, own, 0, nullptr, c.tcc.PARTA + (size_t)((MAX_RB + r) * CL + k.crank) * ActorL::SMALL);
                }
            }
            grid_sync(k);
            stamp(k, 15);
            // =============================== phase D: actor Adam, Polyak, scalars ===============================
            // the Q-filter weight from the published counts (HIRL.py:292-308), identical in every thread
            float wf, omw, cnt = 0.f;
            {
                double w = (double)st.bc_weight_now;
                if (soft) {
                    for (int r = 0; r < nrb; ++r) cnt += __ldcg(c.tcc.CNT + 2 * r);
                    const float mean = st.force_soft_weight >= 0.f ? st.force_soft_weight : cnt / (float)B;
                    w = (double)mean + (double)st.bc_warm_up;
                }
                if (w > 1.0) w = 1.0;
                wf = (float)w;
                omw = (float)(1.0 - w);
            }
            adam_tcc<ActorL, 2 * MAX_RB, true, 3>(k, grank, gsize, 1, c.actor, c.g_actor, c.tcc.GW2PA, c.m_actor, c.v_actor, c.tcc.PARTA, 2 * MAX_RB * CL, 2 * MAX_RB,
                             st.actor_neg_step, st.actor_bc2_sqrt, c.tcc.img_actor, c.tcc.imgt_actor, c.tactor, c.tcc.img_tactor, do_polyak, tau, omt,
                             omw, wf * c.loss_lambda);
            if (do_polyak) polyak_tcc<CriticL>(grank, gsize, 2, c.critic, c.tcritic, c.tcc.img_tcritic, tau, omt);
            if (grank == gsize - 1 && threadIdx.x == 0) {
                float sumq = 0.f, bsq = 0.f, fsq = 0.f;
                for (int r = 0; r < nrb; ++r) {
                    sumq += __ldcg(c.tcc.CNT + 2 * r + 1);
                    bsq += __ldcg(c.tcc.LPT + 8 + 2 * r); fsq += __ldcg(c.tcc.LPT + 8 + 2 * r + 1);
                }
                const float rl_loss = -(sumq / (float)B);                                                     // HIRL.py:297
                const float bc_loss = __fmul_rn(bsq / (float)(A * B), c.loss_lambda);                          // HIRL.py:311
                c.carry[1] = __fadd_rn(__fmul_rn(bc_loss, wf), __fmul_rn(rl_loss, omw));                       // HIRL.py:321
                c.carry[2] = bc_loss;
                c.carry[3] = rl_loss;
                c.carry[4] = fsq / (float)B;                                                                   // HIRL.py:319 (x lambda on the host)
                c.carry[5] = wf;
                c.carry[6] = soft ? cnt : 0.f;
            }
            // the next call's phase A reads nothing that phase D writes unless the targets moved (Polyak): no grid barrier then
            if (do_polyak || step + 1 == Lc.k_steps || c.steps[step + 1].do_actor != 0) grid_sync(k);
            else __syncthreads();
            stamp(k, 16);
        }
        if (grank == gsize - 1 && threadIdx.x < HIRL_OUT_SCALARS) c.out_scalars[(size_t)step * HIRL_OUT_SCALARS + threadIdx.x] = c.carry[threadIdx.x];
    }
    if (grank == 0 && threadIdx.x == 0) *c.tcc.gen_ctr = k.pgen + (unsigned)Lc.k_steps;
    tc::tc_fence_before();
    cluster_sync();                                                    // nobody leaves while a peer may still push into its shared memory
    if (threadIdx.x < 32) tc::tmem_dealloc(k.tm);
}

}  // namespace tcc
}  // namespace hirl
