// cluster_update.cuh -- the cluster-resident fused HIRL update (reference hirl/agents/HIRL.py:221-334).
//
// Round-1's kernel cut a learn() call into 16 stages separated by grid-wide barriers and bounced every activation
// through L2 between them (profiles/r1_stage_timeline_v10_ranks.txt: 11.8 % of the FP32 roofline, barrier- and
// latency-bound).  Here a thread-block CLUSTER of 8 CTAs owns a block of RB = 16 (or 18) batch rows through all layers:
//
//   phase A  per row block, cluster-local: gather -> first layers -> [TA, A, AB, BC] second layers -> LayerNorm2 + heads
//            (row statistics completed across the cluster through DSMEM: st.async + mbarrier, no cluster barrier)
//            -> a' -> target critics -> TD target, critic heads, loss, backward through head / LayerNorm2 -> dz2
//            -> dH1 (partial over each CTA's weight slice, reduce-scattered through DSMEM) -> LayerNorm1 backward
//   ---- grid barrier (the weight gradient needs every row) ----
//   phase B  dW2 tiles with Adam fused on the finished tile; first-layer / small-parameter gradients + Adam
//   ---- grid barrier ----                                                    (critic-only calls end here)
//   phase C  Q1(s,pi), Q1(s,a_bc) with the updated critic, soft count, d pi, actor backward (both branches)
//   ---- grid barrier ----
//   phase D  actor dW2 tiles (RL and BC branch combined with the Q-filter weight, HIRL.py:303-321) + Adam + Polyak
//   ---- grid barrier ----
//
// 4 grid barriers per actor call (2 per critic-only call) instead of 16 (8).  CTA r of a cluster always works on the
// output features 64 r .. 64 r + 63 of the 512-wide layers: its slice of W2 (65 KB, one cp.async.bulk from a padded
// image of the matrix) is streamed into shared memory while the previous pass computes; the slice's LayerNorm / head
// parameters are cached in shared memory once per phase.  The state of a pass (pre-activations, xhat, head outputs)
// lives in shared memory, so GEMMs and epilogues are LOOPS over passes: the first version inlined every pass and was
// 1 MB of straight-line SASS that missed the instruction cache on every pass (profiles/README.md, round 2).
// FP32 FMA throughout.  Everything is reduced in a fixed order: results are run-to-run deterministic.
#pragma once
#include "cluster_gemm.cuh"
#include "update.cuh"

namespace hirl {
namespace cl {

constexpr int MAX_CL = 16;                 // clusters per launch (B200: 15 or 16 co-resident clusters of 8)
constexpr int PC_WORDS = 2 * 4 * H2 + 32;  // per-cluster partial sums of the critic's small second-layer parameters
constexpr int PA_WORDS = 2 * 7 * H2 + 32;  // same for the actor, per branch
constexpr int LPW = 16;                    // per-cluster loss / count words
constexpr int PROF_WORDS = 40;

struct ClusterLaunch {
    AgentCtx agent;
    unsigned *bar_flags;   // [MAX_CL][32] cluster flags of the grid barrier (generations keep counting across launches)
    int k_steps, ncl;
    long long *prof;       // [PROF_STEPS][PROF_WORDS] SM-clock stamps of CTA 0, may be null
};

__device__ __forceinline__ float sum8(float v) {   // lanes l, l^1, l^2, l^4: fixed tree, identical on every lane
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// N independent butterfly reductions interleaved (one dependent shuffle chain per value would expose 5 shuffle
// latencies per value; interleaved the chains hide each other)
template <int N>
__device__ __forceinline__ void warp_sum_n(float (&v)[N]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}
template <int N>
__device__ __forceinline__ void sum8_n(float (&v)[N]) {
#pragma unroll
    for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
}

// parameter cache: per network the slice of LayerNorm2 / head parameters of this CTA's 64 columns
//   [g2 64][be2 64][w3 nout x 64][b3 4]
constexpr int PCA = 6 * NS + 4, PCC = 3 * NS + 4;
enum { N_TA = 0, N_A, N_BC, N_C1, N_C2, N_T1, N_T2 };
__device__ __forceinline__ constexpr int pc_off(int net) { return net < 3 ? net * PCA : 3 * PCA + (net - 3) * PCC; }
constexpr int PC_FLOATS = 3 * PCA + 4 * PCC;

// ---------------------------------------------------------------------------------------------------------------
template <int RI>
struct Ker {
    static constexpr int RB = 2 * RI, RPW = (RB + 7) / 8, RIP = (RI + 3) & ~3, XS = 8 * RB, RL = RedLd<RI>::v;
    static constexpr int RED = 4 * RB * RL > RB * H1 ? 4 * RB * RL : RB * H1;
    static constexpr int DZS1 = NS * 2 * RIP, ROWS = 2 * 3 * H1;
    static constexpr int UNI = 4 * RB * NS > 2 * DZS1 + ROWS ? 4 * RB * NS : 2 * DZS1 + ROWS;
    // shared-memory carve-up (floats).  U is a union: forward pass state ZB[4][RB][64] | dz2 copies + reduce-scatter rows
    static constexpr int O_W0 = 0, O_A0 = 2 * W_FLOATS, O_RED = O_A0 + 2 * RB * WLD, O_U = O_RED + RED, O_DZS = O_U, O_ROWS = O_U + 2 * DZS1,
                         O_XB = O_U + UNI, O_XROW = O_XB + 2 * CL * XS, O_XBROW = O_XROW + RB * 32, O_OUTS = O_XBROW + RB * 20, O_RS = O_OUTS + 4 * RB * 4,
                         O_DOUT = O_RS + 4 * RB, O_AP = O_DOUT + 2 * RB * 4, O_MISC = O_AP + RB * 4, O_DESC = O_MISC + 64, O_PC = O_DESC + 128, TOTAL_A = O_PC + PC_FLOATS;
    static constexpr int TOTAL_B = 256 * NS + 256 * KS + 4 * KS * DLD;   // phase B/D: Dz, Hs, reduction
    static constexpr int TOTAL = TOTAL_A > TOTAL_B ? TOTAL_A : TOTAL_B;
    static constexpr size_t SMEM_BYTES = (size_t)TOTAL * 4;
    static_assert(2 * RB * WLD >= RB * H1, "the second dH1 partial buffer aliases the activation buffers");
    static_assert(RED >= 8 * 7 * NS, "column-sum staging aliases the reduction buffer");
    static_assert(TOTAL_B <= O_A0, "phases B / D only overwrite the weight-slice buffers");
    static_assert(SMEM_BYTES + 1200 <= 232448, "shared memory budget (227 KB per CTA)");

    struct Ctl {
        float *sm;
        uint64_t *wbar, *abar, *xbar;
        int crank, cid, ncl;
        unsigned wi, wo, ai, ao;   // weight-slice / activation-block loads issued and consumed (buffer = index & 1)
        unsigned xr;               // exchange rounds done (parity = xr & 1)
        unsigned tph;              // completed phases of the dW2-tile operand barrier
        unsigned gen;
        unsigned *flags;
        long long *pr;             // profiling stamps of this step (CTA 0, thread 0) or null
    };
    static __device__ __forceinline__ void stamp(const Ctl &k, int id) { if (k.pr) k.pr[id] = clock64(); }
    static __device__ __forceinline__ void gstamp(const Ctl &k, long long *base, int phase) {   // every cluster's CTA 0: phase end on the global timer
        if (base && k.crank == 0 && threadIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            base[phase * MAX_CL + k.cid] = (long long)t;
        }
    }

    // ---- grid barrier: cluster barrier, one flag per cluster, cluster barrier --------------------------------
    static __device__ __forceinline__ void grid_sync(Ctl &k) {
        cluster_sync_all();
        k.gen += 1u;
        if (k.crank == 0 && threadIdx.x < 32) {
            if (threadIdx.x == 0) st_release_u32(k.flags + k.cid * FLAG_STRIDE, k.gen);
            bool pending = true;
            const long long t0 = clock64();
            while (pending) {
                pending = (int)threadIdx.x < k.ncl && (int)(ld_relaxed_u32(k.flags + threadIdx.x * FLAG_STRIDE) - k.gen) < 0;
                pending = __any_sync(0xffffffffu, pending);
                if (pending && clock64() - t0 > 8000000000LL) {   // ~4 s: a lost cluster must fail loudly, not hang the GPU
                    if (threadIdx.x == 0) printf("hirl cluster kernel: grid barrier timed out (cluster %d gen %u)\n", k.cid, k.gen);
                    __trap();
                }
            }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        cluster_sync_all();
    }

    // ---- operand streaming -------------------------------------------------------------------------------------
    static __device__ __forceinline__ void ldw(Ctl &k, const float *w2p) {   // this CTA's 64-row slice of a padded W2 image
        if (threadIdx.x == 0) {
            uint64_t *bar = k.wbar + (k.wi & 1u);
            mbar_expect_tx(bar, W_FLOATS * 4);
            bulk_g2s(k.sm + (k.wi & 1u) * W_FLOATS, w2p + (size_t)k.crank * W_FLOATS, W_FLOATS * 4, bar);
        }
        k.wi += 1u;
    }
    static __device__ __forceinline__ void lda(Ctl &k, const float *h1p_rows) {   // RB padded activation rows (contiguous)
        if (threadIdx.x == 0) {
            uint64_t *bar = k.abar + (k.ai & 1u);
            mbar_expect_tx(bar, RB * WLD * 4);
            bulk_g2s(k.sm + O_A0 + (k.ai & 1u) * RB * WLD, h1p_rows, RB * WLD * 4, bar);
        }
        k.ai += 1u;
    }
    static __device__ __forceinline__ const float *wait_w(const Ctl &k) {
        mbar_wait(k.wbar + (k.wo & 1u), (k.wo >> 1) & 1u, 1);
        return k.sm + (k.wo & 1u) * W_FLOATS;
    }

    // A GEMM pass: which padded weight image, which activation pass, which bias, where the result goes.
    // Descriptor tables (GP, L1D) are built by thread 0 in SHARED memory: as local arrays they lived on the stack, and
    // every cluster-scope fence invalidates L1, so each descriptor read was an L2 round trip on the critical path.
    struct GP { const float *w2p; const float *b2; int a_pass, zslot; };
    static __device__ __forceinline__ GP *gp_table(const Ctl &k) { return reinterpret_cast<GP *>(k.sm + O_DESC); }
    // Runs n forward passes back to back.  The caller has issued the weight load of pass 0; every other operand is
    // requested one pass ahead (while the previous pass computes).  next_w: first weight slice of the next segment.
    // Results (+ bias) land in ZB[zslot] in the epilogue mapping.
    static __device__ __forceinline__ void run_segment(Ctl &k, const AgentCtx &c, const GP *gp, int n, int b0, const float *next_w) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const size_t rowoff = (size_t)b0 * WLD, pstride = (size_t)c.Bpad * WLD;
        lda(k, c.sc.H1P + gp[0].a_pass * pstride + rowoff);
        for (int i = 0; i < n; ++i) {
            const bool last_use = i + 1 == n || gp[i + 1].w2p != gp[i].w2p;
            if (i + 1 < n) {
                if (last_use) ldw(k, gp[i + 1].w2p);
                lda(k, c.sc.H1P + gp[i + 1].a_pass * pstride + rowoff);
            } else if (next_w) {
                ldw(k, next_w);
            }
            const float *b2 = gp[i].b2 + NS * k.crank + lane;
            float b20, b21;   // requested before the waits (volatile: not sunk behind the GEMM), consumed after it
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(b20) : "l"(b2));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(b21) : "l"(b2 + 32));
            const float *ws = wait_w(k);
            mbar_wait(k.abar + (k.ao & 1u), (k.ao >> 1) & 1u, 2);
            fwd_pass<RI>(k.sm + O_A0 + (k.ao & 1u) * RB * WLD, ws, k.sm + O_RED);
            k.ao += 1u;
            if (last_use) k.wo += 1u;
            float *zb = k.sm + O_U + gp[i].zslot * RB * NS;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                if (m >= RB) continue;
                zb[m * NS + lane] = fwd_value<RI>(k.sm + O_RED, m, lane) + b20;
                zb[m * NS + lane + 32] = fwd_value<RI>(k.sm + O_RED, m, lane + 32) + b21;
            }
        }
    }

    // ---- DSMEM exchange rounds ---------------------------------------------------------------------------------
    static __device__ __forceinline__ void x_begin(Ctl &k, int values_expected) {
        __syncthreads();   // everybody has consumed the buffer of two rounds ago; shared state written before is visible
        if (threadIdx.x == 0) mbar_expect_tx(k.xbar + (k.xr & 1u), (uint32_t)values_expected * 4u);
    }
    // lanes 0..7 of the calling warp store v into slot `slot` of this CTA's row in every peer's buffer
    static __device__ __forceinline__ void x_post(const Ctl &k, int slot, float v) {
        const int j = threadIdx.x & 31;
        if (j < CL) {
            const uint32_t par = k.xr & 1u;
            st_async_f32(mapa(smem_u32(k.sm + O_XB + (par * CL + k.crank) * XS + slot), j), v, mapa(smem_u32(k.xbar + par), j));
        }
    }
    static __device__ __forceinline__ const float *x_wait(Ctl &k) {   // returns the round's buffer [CL][XS]
        const uint32_t par = k.xr & 1u;
        mbar_wait(k.xbar + par, (k.xr >> 1) & 1u, 3);
        k.xr += 1u;
        return k.sm + O_XB + par * CL * XS;
    }

    // ---- parameter cache -----------------------------------------------------------------------------------------
    template <class L>
    static __device__ __forceinline__ void fill_pc(const Ctl &k, int net, const float *tw) {
        float *dst = k.sm + O_PC + pc_off(net);
        const int n0 = NS * k.crank;
        for (int i = threadIdx.x; i < (2 + L::nout) * NS + 4; i += CTHREADS) {
            float v;
            if (i < NS) v = tw[L::G2 + n0 + i];
            else if (i < 2 * NS) v = tw[L::BE2 + n0 + i - NS];
            else if (i < (2 + L::nout) * NS) { const int o = (i - 2 * NS) / NS, j = (i - 2 * NS) & (NS - 1); v = tw[L::W3 + o * H2 + n0 + j]; }
            else { const int o = i - (2 + L::nout) * NS; v = o < L::nout ? tw[L::B3 + o] : 0.f; }
            dst[i] = v;
        }
    }

    // ---- epilogues in the mapping "warp w: rows m = w + 8 r, lane l: columns l, l + 32 of the slice" -------------------
    struct EP { int net, nout, zslot; };   // parameter-cache entry, head width, pass-state slot
    // LayerNorm2 statistics of NP passes in one exchange round: z -> xhat in place, rstd -> RS[zslot][m]
    template <int NP>
    static __device__ __forceinline__ void stats_round_n(Ctl &k, const EP *ep) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        constexpr int NV = NP * RPW;
        float z0[NV], z1[NV], mean[NV], m2[NV];
        x_begin(k, NP * 2 * RB * CL);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *zb = k.sm + O_U + ep[p].zslot * RB * NS;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = p * RPW + r;
                z0[i] = m < RB ? zb[m * NS + lane] : 0.f;
                z1[i] = m < RB ? zb[m * NS + lane + 32] : 0.f;
                mean[i] = z0[i] + z1[i];
            }
        }
        warp_sum_n<NV>(mean);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            mean[i] *= 1.f / NS;
            const float d0 = z0[i] - mean[i], d1 = z1[i] - mean[i];
            m2[i] = fmaf(d0, d0, d1 * d1);
        }
        warp_sum_n<NV>(m2);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = p * RPW + r;
                if (m >= RB) continue;
                x_post(k, (p * RB + m) * 2, mean[i]);
                x_post(k, (p * RB + m) * 2 + 1, m2[i]);
            }
        const float *xb = x_wait(k);
        // Chan's parallel variance over the 8 slices of 64 columns
        float mi[NV], m2i[NV], mu[NV];
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = p * RPW + r;
                mi[i] = m < RB ? xb[(lane & 7) * XS + (p * RB + m) * 2] : 0.f;
                m2i[i] = m < RB ? xb[(lane & 7) * XS + (p * RB + m) * 2 + 1] : 0.f;
                mu[i] = mi[i];
            }
        sum8_n<NV>(mu);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            mu[i] *= 1.f / CL;
            const float d = mi[i] - mu[i];
            m2i[i] = fmaf((float)NS * d, d, m2i[i]);
        }
        sum8_n<NV>(m2i);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float *zb = k.sm + O_U + ep[p].zslot * RB * NS;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = p * RPW + r;
                if (m >= RB) continue;
                const float rstd = 1.f / sqrtf(m2i[i] * (1.f / H2) + LN_EPS);
                zb[m * NS + lane] = (z0[i] - mu[i]) * rstd;
                zb[m * NS + lane + 32] = (z1[i] - mu[i]) * rstd;
                if (lane == 0) k.sm[O_RS + ep[p].zslot * RB + m] = rstd;
            }
        }
    }
    static __device__ __forceinline__ void stats_round(Ctl &k, const EP *ep, int np) {
        if (np == 4) stats_round_n<4>(k, ep);
        else if (np == 2) stats_round_n<2>(k, ep);
        else stats_round_n<1>(k, ep);
    }
    // heads of NP passes (ep[0..NP)) of width NOUT in one exchange round: OUTS[zslot][m][o]; ends with a CTA barrier
    template <int NP, int NOUT>
    static __device__ __forceinline__ void head_round_n(Ctl &k, const EP *ep, float slope) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        constexpr int NV = NP * RPW * NOUT;
        float d[NV];
        x_begin(k, NP * NOUT * RB * CL);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *zb = k.sm + O_U + ep[p].zslot * RB * NS, *pc = k.sm + O_PC + pc_off(ep[p].net);
            const float g0 = pc[lane], g1 = pc[lane + 32], e0 = pc[NS + lane], e1 = pc[NS + lane + 32];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                const float h0 = m < RB ? act_f(fmaf(zb[m * NS + lane], g0, e0), slope) : 0.f, h1 = m < RB ? act_f(fmaf(zb[m * NS + lane + 32], g1, e1), slope) : 0.f;
#pragma unroll
                for (int o = 0; o < NOUT; ++o) d[(p * RPW + r) * NOUT + o] = fmaf(h0, pc[(2 + o) * NS + lane], h1 * pc[(2 + o) * NS + lane + 32]);
            }
        }
        warp_sum_n<NV>(d);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                if (m >= RB) continue;
#pragma unroll
                for (int o = 0; o < NOUT; ++o) x_post(k, (p * RB + m) * NOUT + o, d[(p * RPW + r) * NOUT + o]);
            }
        const float *xb = x_wait(k);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
#pragma unroll
                for (int o = 0; o < NOUT; ++o) d[(p * RPW + r) * NOUT + o] = m < RB ? xb[(lane & 7) * XS + (p * RB + m) * NOUT + o] : 0.f;
            }
        sum8_n<NV>(d);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *pc = k.sm + O_PC + pc_off(ep[p].net);
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                if (m >= RB || lane != 0) continue;
#pragma unroll
                for (int o = 0; o < NOUT; ++o) k.sm[O_OUTS + (ep[p].zslot * RB + m) * 4 + o] = d[(p * RPW + r) * NOUT + o] + pc[(2 + NOUT) * NS + o];
            }
        }
        __syncthreads();
    }
    // heads of passes [p0, p1) (all of the same width)
    static __device__ __forceinline__ void head_round(Ctl &k, const EP *ep, int p0, int p1, float slope) {
        const int np = p1 - p0;
        if (ep[p0].nout == A) {
            if (np == 2) head_round_n<2, A>(k, ep + p0, slope);
            else head_round_n<1, A>(k, ep + p0, slope);
        } else {
            if (np == 4) head_round_n<4, 1>(k, ep + p0, slope);
            else if (np == 2) head_round_n<2, 1>(k, ep + p0, slope);
            else head_round_n<1, 1>(k, ep + p0, slope);
        }
    }

    // Backward through head + LayerNorm2 of NP passes (upstream gradients in DOUT[p][m][4]), column sums of the small
    // second-layer parameters over the block's rows, dz2 -> shared copy (+ slice-major global scratch for dW2).
    //   cacc[p][u]: this thread's running column sums (item = threadIdx.x + 256 u of [q][64], q = db2, dg2, dbe2, dW3[o])
    //   db3[p]: thread o < nout accumulates sum_rows dout[o]
    template <int NP>
    static __device__ __forceinline__ void ln2_backward(Ctl &k, const AgentCtx &c, const EP *ep, const int *gslot, int b0, float slope, float (&cacc)[2][2],
                                                        float (&db3)[2]) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float dy[NP][RPW][2], dz[NP][RPW][2], sx[NP * RPW * 2];
        x_begin(k, NP * 2 * RB * CL);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *zb = k.sm + O_U + ep[p].zslot * RB * NS, *pc = k.sm + O_PC + pc_off(ep[p].net), *dout = k.sm + O_DOUT + p * RB * 4;
            const int nout = ep[p].nout;
            const float g0 = pc[lane], g1 = pc[lane + 32], e0 = pc[NS + lane], e1 = pc[NS + lane + 32];
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = (p * RPW + r) * 2;
                dy[p][r][0] = dy[p][r][1] = dz[p][r][0] = dz[p][r][1] = 0.f;
                sx[i] = sx[i + 1] = 0.f;
                if (m >= RB) continue;
                const float x0 = zb[m * NS + lane], x1 = zb[m * NS + lane + 32];
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < nout) {
                        d0 = fmaf(dout[m * 4 + o], pc[(2 + o) * NS + lane], d0);
                        d1 = fmaf(dout[m * 4 + o], pc[(2 + o) * NS + lane + 32], d1);
                    }
                d0 *= act_g(fmaf(x0, g0, e0), slope);
                d1 *= act_g(fmaf(x1, g1, e1), slope);
                dy[p][r][0] = d0; dy[p][r][1] = d1;
                dz[p][r][0] = d0 * g0; dz[p][r][1] = d1 * g1;   // d xhat
                sx[i] = dz[p][r][0] + dz[p][r][1];
                sx[i + 1] = fmaf(dz[p][r][0], x0, dz[p][r][1] * x1);
            }
        }
        warp_sum_n<NP * RPW * 2>(sx);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = (p * RPW + r) * 2;
                if (m >= RB) continue;
                x_post(k, (p * RB + m) * 2, sx[i]);
                x_post(k, (p * RB + m) * 2 + 1, sx[i + 1]);
            }
        const float *xb = x_wait(k);
#pragma unroll
        for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = (p * RPW + r) * 2;
                sx[i] = m < RB ? xb[(lane & 7) * XS + (p * RB + m) * 2] : 0.f;
                sx[i + 1] = m < RB ? xb[(lane & 7) * XS + (p * RB + m) * 2 + 1] : 0.f;
            }
        sum8_n<NP * RPW * 2>(sx);
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *zb = k.sm + O_U + ep[p].zslot * RB * NS;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, i = (p * RPW + r) * 2;
                if (m >= RB) continue;
                const float s1 = sx[i] * (1.f / H2), s2 = sx[i + 1] * (1.f / H2);
                const float rstd = k.sm[O_RS + ep[p].zslot * RB + m];
                dz[p][r][0] = rstd * (dz[p][r][0] - s1 - zb[m * NS + lane] * s2);
                dz[p][r][1] = rstd * (dz[p][r][1] - s1 - zb[m * NS + lane + 32] * s2);
            }
        }
        // column sums, one pass at a time: cs[warp][q][64] staged in the reduction buffer
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const float *zb = k.sm + O_U + ep[p].zslot * RB * NS, *pc = k.sm + O_PC + pc_off(ep[p].net), *dout = k.sm + O_DOUT + p * RB * 4;
            const int nout = ep[p].nout, nq = 3 + nout;
            const float g0 = pc[lane], g1 = pc[lane + 32], e0 = pc[NS + lane], e1 = pc[NS + lane + 32];
            float *cs = k.sm + O_RED + warp * nq * NS;
            float s[7][2];
#pragma unroll
            for (int q = 0; q < 7; ++q) s[q][0] = s[q][1] = 0.f;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                if (m >= RB) continue;
                const float x0 = zb[m * NS + lane], x1 = zb[m * NS + lane + 32];
                const float h0 = act_f(fmaf(x0, g0, e0), slope), h1 = act_f(fmaf(x1, g1, e1), slope);
                s[0][0] += dz[p][r][0]; s[0][1] += dz[p][r][1];
                s[1][0] = fmaf(dy[p][r][0], x0, s[1][0]); s[1][1] = fmaf(dy[p][r][1], x1, s[1][1]);
                s[2][0] += dy[p][r][0]; s[2][1] += dy[p][r][1];
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < nout) { s[3 + o][0] = fmaf(dout[m * 4 + o], h0, s[3 + o][0]); s[3 + o][1] = fmaf(dout[m * 4 + o], h1, s[3 + o][1]); }
            }
            __syncthreads();   // previous user of the staging buffer is done
#pragma unroll
            for (int q = 0; q < 7; ++q)
                if (q < nq) { cs[q * NS + lane] = s[q][0]; cs[q * NS + lane + 32] = s[q][1]; }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int item = threadIdx.x + u * CTHREADS;
                if (item < nq * NS) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) t += k.sm[O_RED + w * nq * NS + item];
                    cacc[p][u] += t;
                }
            }
            if ((int)threadIdx.x < nout) {
                float t = 0.f;
                for (int m = 0; m < RB; ++m) t += dout[m * 4 + threadIdx.x];
                db3[p] += t;
            }
        }
        __syncthreads();   // every read of the pass state (ZB) is done: the dz2 copies may overwrite it (shared union)
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            float *dzs = k.sm + O_DZS + p * DZS1;
            float *g = gslot[p] >= 0 ? c.sc.DZ2S + ((size_t)(gslot[p] * CL + k.crank) * c.Bpad) * NS : nullptr;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r;
                if (m >= RB) continue;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int n = lane + 32 * cc;
                    dzs[(n * 2 + (m & 1)) * RIP + (m >> 1)] = dz[p][r][cc];
                    if (g && b0 + m < c.B) g[(size_t)(b0 + m) * NS + n] = dz[p][r][cc];
                }
            }
        }
    }

    // dH1 of one or two passes (partials -> reduce-scatter); results in ROWS[which][q][256]
    static __device__ __forceinline__ void dh1(Ctl &k, const float *w0, const float *w1, bool two) {
        __syncthreads();   // dz2 copies complete
        bwd_pass<RI>(k.sm + O_DZS, w0, k.sm + O_RED);
        if (two) bwd_pass<RI>(k.sm + O_DZS + DZS1, w1, k.sm + O_A0);
        cluster_sync_all();
        reduce_scatter_rows<RB>(k.sm + O_RED, k.crank, k.sm + O_ROWS);
        if (two) reduce_scatter_rows<RB>(k.sm + O_A0, k.crank, k.sm + O_ROWS + 3 * H1);
        __syncthreads();
        // the partial buffers are overwritten only after a later cluster-wide barrier (next first-layer hand-over or
        // the grid barrier), i.e. after every peer has finished these remote reads
    }

    // ---- gather (buffer.py:38-48, HIRL.py:229-233, 245-251) into shared rows --------------------------------------
    static __device__ __forceinline__ void gather_rows(const Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step, int b0, bool replay, bool bc) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B;
        const int32_t *ion = c.idx_online + (size_t)step * B, *iex = c.idx_expert + (size_t)step * B, *ibc = c.idx_bc + (size_t)step * B;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r, b = b0 + m;
            if (m >= RB) continue;
            if (replay) {
                float v = 0.f;
                if (b < B) {
                    const float *src = b < st.n_online ? c.online + (size_t)__ldg(ion + b) * ROW : c.expert + (size_t)__ldg(iex + (b - st.n_online)) * ROW;
                    v = __ldg(src + lane);
                    if (k.crank == 0) c.sc.X[(size_t)b * ROW + lane] = v;
                }
                k.sm[O_XROW + m * 32 + lane] = v;
            }
            if (bc && lane < BCW) {
                float v = 0.f;
                if (b < B) {
                    v = __ldg(c.bctab + (size_t)__ldg(ibc + b) * BCW + lane);
                    if (k.crank == 0) c.sc.XB[(size_t)b * BCW + lane] = v;
                }
                k.sm[O_XBROW + m * BCW + lane] = v;
            }
        }
    }

    // First layers (csrc/update.cuh: l1_row) of the rows this CTA owns (m = crank + 8 q), one warp per (pass, row).
    //   actor-type net: state = words [xoff, xoff + 13) of row m of xsrc;
    //   critic tower:   state as above, action = asrc[m][4] (or the 4 words after the state when asrc is null)
    struct L1D { const float *tw; const float *xsrc; const float *asrc; int xld, xoff, pass, critic; };
    static __device__ __forceinline__ L1D *l1_table(const Ctl &k) { return reinterpret_cast<L1D *>(k.sm + O_DESC + 32); }
    static __device__ __noinline__ void l1_rows(const AgentCtx &c, const L1D *d, int nd, int crank, int b0, float slope) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int inst = warp; inst < nd * RPW; inst += WARPS) {
            const int pi = inst / RPW, q = inst - pi * RPW, m = crank + 8 * q, b = b0 + m;
            if (m >= RB) continue;
            const L1D &e = d[pi];
            const int slot = bwd_slot(e.pass);
            const bool keep = slot >= 0 && b < c.B;
            float *h1 = c.sc.H1P + ((size_t)e.pass * c.Bpad + b) * WLD;
            float *xh = keep ? XH1p(c, slot) + (size_t)b * H1 : nullptr, *rs = keep ? RS1p(c, slot) + b : nullptr;
            const float *xs = e.xsrc + m * e.xld + e.xoff;
            if (e.critic) {
                const float *as = e.asrc ? e.asrc + m * A : xs + S;
                float x[SA];
#pragma unroll
                for (int kk = 0; kk < S; ++kk) x[kk] = xs[kk];
#pragma unroll
                for (int j = 0; j < A; ++j) x[S + j] = as[j];
                l1_row<CriticL>(e.tw, x, slope, lane, h1, xh, rs);
            } else {
                float x[S];
#pragma unroll
                for (int kk = 0; kk < S; ++kk) x[kk] = xs[kk];
                l1_row<ActorL>(e.tw, x, slope, lane, h1, xh, rs);
            }
        }
    }

    // LayerNorm1 backward of the rows this CTA owns; dh1 rows come from the reduce-scatter buffer.  Writes dz1 / dy1 for
    // the first-layer gradient pass; optionally d(action) (HIRL.py:296: Q1's gradient w.r.t. pi).
    template <class L>
    static __device__ __forceinline__ void ln1_rows(const Ctl &k, const AgentCtx &c, const float *tw, int which, int slot, int b0, float slope, int q, bool store,
                                                    float *gpi4) {
        const int lane = threadIdx.x & 31, m = k.crank + 8 * q, b = b0 + m;
        if (m >= RB || b >= c.B) {
            if (gpi4) gpi4[0] = gpi4[1] = gpi4[2] = gpi4[3] = 0.f;
            return;
        }
        const float *dsrc = k.sm + O_ROWS + (which * 3 + q) * H1;
        const float *xs = XH1p(c, slot) + (size_t)b * H1;
        float dh[8], xh[8], dz[8], dy[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { dh[j] = dsrc[lane + 32 * j]; xh[j] = xs[lane + 32 * j]; }
        ln1_backward<L>(tw, dh, xh, RS1p(c, slot)[b], slope, lane, dz, dy);
        if (store) {
            float *dz1 = c.sc.DZ1 + ((size_t)slot * c.B + b) * H1, *dy1 = c.sc.DY1 + ((size_t)slot * c.B + b) * H1;
#pragma unroll
            for (int j = 0; j < 8; ++j) { dz1[lane + 32 * j] = dz[j]; dy1[lane + 32 * j] = dy[j]; }
        }
        if (gpi4) {
#pragma unroll
            for (int j2 = 0; j2 < A; ++j2) {
                float p = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) p = fmaf(dz[j], tw[L::W1T + (S + j2) * H1 + lane + 32 * j], p);
                gpi4[j2] = warp_sum(p);
            }
        }
    }

    // per-cluster scalar partial: sum over the CTA's warps of a lane-0 value, written by CTA 0 of the cluster
    static __device__ __forceinline__ void flush_scalar(const Ctl &k, const AgentCtx &c, int word, float lane0_value) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        __syncthreads();
        if (lane == 0) k.sm[O_MISC + 32 + warp] = lane0_value;
        __syncthreads();
        if (threadIdx.x == 0 && k.crank == 0) {
            float s = 0.f;
            for (int w = 0; w < WARPS; ++w) s += k.sm[O_MISC + 32 + w];
            c.sc.LP[k.cid * LPW + word] = s;
        }
    }
    // keeps the forward state of a pass for the actor backward in phase C
    static __device__ __forceinline__ void save_pass(const Ctl &k, const AgentCtx &c, int zslot, int which, int b0) {
        const float *zb = k.sm + O_U + zslot * RB * NS;
        for (int i = threadIdx.x; i < RB * NS; i += CTHREADS) {
            const int m = i >> 6, n = i & 63, b = b0 + m;
            if (b < c.B) c.sc.XH2[((size_t)which * c.Bpad + b) * H2 + NS * k.crank + n] = zb[i];
        }
        if (k.crank == 0 && threadIdx.x < RB && b0 + (int)threadIdx.x < c.B)
            c.sc.RS2[(size_t)which * c.Bpad + b0 + threadIdx.x] = k.sm[O_RS + zslot * RB + threadIdx.x];
    }
    static __device__ __forceinline__ void load_pass(const Ctl &k, const AgentCtx &c, int zslot, int which, int b0) {
        float *zb = k.sm + O_U + zslot * RB * NS;
        for (int i = threadIdx.x; i < RB * NS; i += CTHREADS) {
            const int m = i >> 6, n = i & 63, b = b0 + m;
            zb[i] = b < c.B ? c.sc.XH2[((size_t)which * c.Bpad + b) * H2 + NS * k.crank + n] : 0.f;
        }
        if (threadIdx.x < RB) k.sm[O_RS + zslot * RB + threadIdx.x] = b0 + (int)threadIdx.x < c.B ? c.sc.RS2[(size_t)which * c.Bpad + b0 + threadIdx.x] : 0.f;
    }

    // ===========================================================================================================
    // phase A (modes 0 / 1): everything of a call that is local to a row block, up to the critic's dz2 / dz1
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_a(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        const bool act = st.do_actor != 0, hirl = st.mode == 0;
        const float *c2 = c.critic + CriticL::SIZE, *t2 = c.tcritic + CriticL::SIZE;
        if (k.cid < nblk) {
            fill_pc<ActorL>(k, N_TA, c.tactor);
            fill_pc<CriticL>(k, N_C1, c.critic); fill_pc<CriticL>(k, N_C2, c2);
            fill_pc<CriticL>(k, N_T1, c.tcritic); fill_pc<CriticL>(k, N_T2, t2);
            if (act) fill_pc<ActorL>(k, N_A, c.actor);
            if (act && hirl) fill_pc<ActorL>(k, N_BC, c.bc);
        }
        float cacc[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, db3[2] = {0.f, 0.f}, loss0 = 0.f, loss1 = 0.f;
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_tactor);
            gather_rows(k, c, st, step, b0, true, act && hirl);
            __syncthreads();
            {   // first layers of the passes whose inputs are known
                L1D *d = l1_table(k);
                const float *xr = k.sm + O_XROW, *xbr = k.sm + O_XBROW;
                const int nd = act ? (hirl ? 6 : 4) : 3;
                if (threadIdx.x == 0) {
                    d[0] = L1D{c.tactor, xr, nullptr, 32, SA, P_TA, 0};
                    d[1] = L1D{c.critic, xr, nullptr, 32, 0, P_C1, 1};
                    d[2] = L1D{c2, xr, nullptr, 32, 0, P_C2, 1};
                    d[3] = L1D{c.actor, xr, nullptr, 32, 0, P_A, 0};
                    d[4] = L1D{c.actor, xbr, nullptr, BCW, 0, P_AB, 0};
                    d[5] = L1D{c.bc, xr, nullptr, 32, 0, P_BC, 0};
                    GP *gp = gp_table(k);
                    gp[0] = GP{c.w2p_tactor, c.tactor + ActorL::B2, P_TA, 0};
                    gp[1] = GP{c.w2p_actor, c.actor + ActorL::B2, P_A, 1};
                    gp[2] = GP{c.w2p_actor, c.actor + ActorL::B2, P_AB, 2};
                    gp[3] = GP{c.w2p_bc, c.bc + ActorL::B2, P_BC, 3};
                }
                __syncthreads();
                l1_rows(c, d, nd, k.crank, b0, slope);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            stamp(k, 1);
            // ---- segment 1: TA [A, AB, BC] -> ZB slots 0..3
            {
                const int n = act ? (hirl ? 4 : 2) : 1;
                run_segment(k, c, gp_table(k), n, b0, c.w2p_tcritic);
                stamp(k, 2);
                const EP ep[4] = {{N_TA, A, 0}, {N_A, A, 1}, {N_A, A, 2}, {N_BC, A, 3}};
                stats_round(k, ep, n);
                head_round(k, ep, 0, n < 2 ? n : 2, slope);
                if (n > 2) head_round(k, ep, 2, n, slope);
            }
            // a' = clamp(targetActor(s') + clamp(noise)) (HIRL.py:264-267); pi, pi_bc, a_bc for phase C
            if (threadIdx.x < RB * A) {
                const int m = threadIdx.x >> 2, j = threadIdx.x & 3, b = b0 + m;
                const float *o = k.sm + O_OUTS + m * 4 + j;
                const float nz = fminf(fmaxf(st.noise[j], -c.noise_clamp), c.noise_clamp);
                const float ap = fminf(fmaxf(tanhf(o[0]) + nz, -1.f), 1.f);
                k.sm[O_AP + threadIdx.x] = ap;
                if (k.crank == 0 && b < B) {
                    c.sc.APRIME[(size_t)b * A + j] = ap;
                    if (act) c.sc.PI[(size_t)b * A + j] = tanhf(o[RB * 4]);
                    if (act && hirl) { c.sc.PIB[(size_t)b * A + j] = tanhf(o[2 * RB * 4]); c.sc.ABC[(size_t)b * A + j] = tanhf(o[3 * RB * 4]); }
                }
            }
            if (act) save_pass(k, c, 1, 0, b0);
            if (act && hirl) save_pass(k, c, 2, 1, b0);
            __syncthreads();
            stamp(k, 3);
            {   // first layers of the target critics on (s', a')  (HIRL.py:270)
                L1D *d = l1_table(k);
                if (threadIdx.x == 0) {
                    d[0] = L1D{c.tcritic, k.sm + O_XROW, k.sm + O_AP, 32, SA, P_TC1, 1};
                    d[1] = L1D{t2, k.sm + O_XROW, k.sm + O_AP, 32, SA, P_TC2, 1};
                    GP *gp = gp_table(k);
                    gp[0] = GP{c.w2p_tcritic, c.tcritic + CriticL::B2, P_TC1, 0};
                    gp[1] = GP{c.w2p_tcritic + (size_t)H2 * WLD, t2 + CriticL::B2, P_TC2, 1};
                    gp[2] = GP{c.w2p_critic, c.critic + CriticL::B2, P_C1, 2};
                    gp[3] = GP{c.w2p_critic + (size_t)H2 * WLD, c2 + CriticL::B2, P_C2, 3};
                }
                __syncthreads();
                l1_rows(c, d, 2, k.crank, b0, slope);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            stamp(k, 4);
            // ---- segment 2: TC1, TC2, C1, C2 -> ZB slots 0..3
            const EP ep[4] = {{N_T1, 1, 0}, {N_T2, 1, 1}, {N_C1, 1, 2}, {N_C2, 1, 3}};
            run_segment(k, c, gp_table(k), 4, b0, nullptr);
            const float *wC2 = k.sm + ((k.wo - 1u) & 1u) * W_FLOATS, *wC1 = k.sm + (k.wo & 1u) * W_FLOATS;   // the last two slices stay resident
            stamp(k, 5);
            stats_round(k, ep, 4);
            head_round(k, ep, 0, 4, slope);
            // TD target, losses, upstream gradients (HIRL.py:270-284)
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, b = b0 + m;
                if (m >= RB || lane != 0) continue;
                float dq1 = 0.f, dq2 = 0.f;
                if (b < B) {
                    const float *o = k.sm + O_OUTS + m * 4;
                    const float t1 = o[0], t2v = o[RB * 4], q1 = o[2 * RB * 4], q2 = o[3 * RB * 4];
                    const float rew = k.sm[O_XROW + m * 32 + 30], d = k.sm[O_XROW + m * 32 + 31];
                    const float tq = fminf(t1, t2v);
                    const float y = __fadd_rn(rew, __fmul_rn(__fmul_rn(c.gamma, tq), __fsub_rn(1.f, d)));   // HIRL.py:274
                    const float d1 = q1 - y, d2 = q2 - y;
                    loss0 = fmaf(d1, d1, loss0);
                    loss1 = fmaf(d2, d2, loss1);
                    dq1 = (2.f / (float)B) * d1;   // mse_loss backward, mean reduction (HIRL.py:284)
                    dq2 = (2.f / (float)B) * d2;
                    if (k.crank == 0) { c.sc.Q[b] = q1; c.sc.Q[(size_t)B + b] = q2; c.sc.TQ[b] = t1; c.sc.TQ[(size_t)B + b] = t2v; c.sc.Y[b] = y; }
                }
                k.sm[O_DOUT + m * 4] = dq1;
                k.sm[O_DOUT + RB * 4 + m * 4] = dq2;
            }
            stamp(k, 6);
            const int gslot[2] = {0, 1};
            ln2_backward<2>(k, c, ep + 2, gslot, b0, slope, cacc, db3);
            stamp(k, 7);
            dh1(k, wC1, wC2, true);
            stamp(k, 8);
            for (int inst = warp; inst < 2 * RPW; inst += WARPS) {
                const int t = inst / RPW, q = inst - t * RPW;
                ln1_rows<CriticL>(k, c, c.critic + t * CriticL::SIZE, t, t, b0, slope, q, true, nullptr);
            }
        }
        // ---- per-cluster partial sums -> global
        if (k.cid < nblk) {
            float *part = c.sc.PART2C + (size_t)k.cid * PC_WORDS;
            const int q = threadIdx.x >> 6, n = NS * k.crank + (threadIdx.x & 63);
            part[(0 * 4 + q) * H2 + n] = cacc[0][0];
            part[(1 * 4 + q) * H2 + n] = cacc[1][0];
            if (k.crank == 0 && threadIdx.x == 0) { part[2 * 4 * H2] = db3[0]; part[2 * 4 * H2 + 1] = db3[1]; }
            flush_scalar(k, c, 0, loss0);
            flush_scalar(k, c, 1, loss1);
        }
    }

    // ---- actor backward for one row block (HIRL.py:295-321; BC.py:177-183), shared by phase C and the BC mode -----
    // branch rl: batch states (pass A, state slot 2, gradient slot 2), upstream d pi from the critic at unit scale (the
    // (1-w) factor is applied in phase D); branch bc: expert states (pass AB, state slot 3, gradient slot 3).
    // Expects the pass states (xhat, rstd) in ZB slots 2 / 3 and d pi of every row in O_AP for the rl branch.
    struct ActAcc { float cacc[2][2], db3[2], bc_sq, fire_sq; };
    static __device__ __forceinline__ void actor_backward(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int b0, bool rl, bool bc, ActAcc &aa) {
        const int warp = threadIdx.x >> 5, B = c.B;
        const float slope = slope_of(c);
        const float bc_scale = (2.f / (float)(A * B)) * (st.mode == 2 ? 1.f : c.loss_lambda);
        const int nb = (rl ? 1 : 0) + (bc ? 1 : 0);
        // upstream gradients at the tanh inputs: DOUT[p][m][j]
        if (threadIdx.x < RB * A) {
            const int m = threadIdx.x >> 2, j = threadIdx.x & 3, b = b0 + m;
            float sq = 0.f, fq = 0.f;
            if (rl) {
                const float pi = b < B ? c.sc.PI[(size_t)b * A + j] : 0.f;
                k.sm[O_DOUT + threadIdx.x] = b < B ? k.sm[O_AP + threadIdx.x] * (1.f - pi * pi) : 0.f;   // tanh backward
            }
            if (bc) {
                const float pi = b < B ? c.sc.PIB[(size_t)b * A + j] : 0.f;
                const float diff = pi - k.sm[O_XBROW + m * BCW + S + j];
                if (b < B) { sq = diff * diff; fq = j == 3 ? sq : 0.f; }
                k.sm[O_DOUT + (rl ? RB * 4 : 0) + threadIdx.x] = b < B ? bc_scale * diff * (1.f - pi * pi) : 0.f;
            }
            aa.bc_sq += sq; aa.fire_sq += fq;
        }
        const EP ep2[2] = {{N_A, A, 2}, {N_A, A, 3}};
        const EP *ep = rl ? ep2 : ep2 + 1;
        const int gs2[2] = {2, 3};
        const int *gslot = rl ? gs2 : gs2 + 1;
        if (nb == 2) {
            ln2_backward<2>(k, c, ep, gslot, b0, slope, aa.cacc, aa.db3);
        } else {
            float ca1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, d1[2] = {0.f, 0.f};
            ln2_backward<1>(k, c, ep, gslot, b0, slope, ca1, d1);
            const int br = rl ? 0 : 1;
            aa.cacc[br][0] += ca1[0][0]; aa.cacc[br][1] += ca1[0][1]; aa.db3[br] += d1[0];
        }
        const float *ws = wait_w(k);   // the actor's slice (requested while the last forward pass ran)
        dh1(k, ws, ws, nb == 2);
        k.wo += 1u;
        for (int inst = warp; inst < nb * RPW; inst += WARPS) {
            const int which = inst / RPW, q = inst - which * RPW;
            const int slot = (which == 0 && rl) ? 2 : 3;
            ln1_rows<ActorL>(k, c, c.actor, which, slot, b0, slope, q, true, nullptr);
        }
    }
    static __device__ __forceinline__ void flush_actor(const Ctl &k, const AgentCtx &c, const ActAcc &aa, bool rl, bool bc) {
        float *part = c.sc.PART2A + (size_t)k.cid * PA_WORDS;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            if (!(br ? bc : rl)) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int item = threadIdx.x + u * CTHREADS;
                if (item < 7 * NS) part[(br * 7 + item / NS) * H2 + NS * k.crank + (item & (NS - 1))] = aa.cacc[br][u];
            }
            if (k.crank == 0 && threadIdx.x < A) part[2 * 7 * H2 + br * 4 + threadIdx.x] = aa.db3[br];
        }
        // bc_sq / fire_sq live on threads < RB*4 (warps 0..2): warp sums, then the per-cluster scalar
        const float s = warp_sum(aa.bc_sq), f = warp_sum(aa.fire_sq);
        flush_scalar(k, c, 4, s);
        flush_scalar(k, c, 5, f);
    }

    // ===========================================================================================================
    // phase C (modes 0 / 1, actor calls): Q1 with the UPDATED critic, soft count, d pi, actor backward
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_c(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        const bool hirl = st.mode == 0, soft = hirl && st.bc_weight_now == 100.f;
        ActAcc aa;
        aa.cacc[0][0] = aa.cacc[0][1] = aa.cacc[1][0] = aa.cacc[1][1] = 0.f;
        aa.db3[0] = aa.db3[1] = 0.f; aa.bc_sq = aa.fire_sq = 0.f;
        float cnt = 0.f, sum_q = 0.f;
        if (k.cid < nblk) fill_pc<CriticL>(k, N_C1, c.critic);   // the critic changed in phase B
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_critic);
            gather_rows(k, c, st, step, b0, true, hirl);
            for (int i = threadIdx.x; i < RB * A * 2; i += CTHREADS) {   // pi, a_bc of the block -> shared (critic inputs)
                const int which = i / (RB * A), e = i - which * RB * A, b = b0 + (e >> 2);
                if (which == 0 || soft) k.sm[O_DOUT + i] = b < B ? (which ? c.sc.ABC : c.sc.PI)[(size_t)b * A + (e & 3)] : 0.f;
            }
            {
                L1D *d = l1_table(k);
                if (threadIdx.x == 0) {
                    d[0] = L1D{c.critic, k.sm + O_XROW, k.sm + O_DOUT, 32, 0, P_QP, 1};
                    d[1] = L1D{c.critic, k.sm + O_XROW, k.sm + O_DOUT + RB * A, 32, 0, P_QS, 1};
                    GP *gp = gp_table(k);
                    gp[0] = GP{c.w2p_critic, c.critic + CriticL::B2, P_QP, 2};
                    gp[1] = GP{c.w2p_critic, c.critic + CriticL::B2, P_QS, 3};
                }
                __syncthreads();
                l1_rows(c, d, soft ? 2 : 1, k.crank, b0, slope);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            const EP ep[2] = {{N_C1, 1, 2}, {N_C1, 1, 3}};
            run_segment(k, c, gp_table(k), soft ? 2 : 1, b0, c.w2p_actor);
            const float *wQ = k.sm + ((k.wo - 1u) & 1u) * W_FLOATS;
            stamp(k, 13);
            stats_round(k, ep, soft ? 2 : 1);
            head_round(k, ep, 0, soft ? 2 : 1, slope);
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, b = b0 + m;
                if (m >= RB || lane != 0) continue;
                float dq = 0.f;
                if (b < B) {
                    const float q = k.sm[O_OUTS + (2 * RB + m) * 4], sq = k.sm[O_OUTS + (3 * RB + m) * 4];
                    dq = -1.f / (float)B;                  // rl_loss = -mean(Q1)  (HIRL.py:297)
                    sum_q += q;
                    if (soft) cnt += sq > q ? 1.f : 0.f;   // HIRL.py:303
                    if (k.crank == 0) { c.sc.RLQ[b] = q; if (soft) c.sc.SOFTQ[b] = sq; }
                }
                k.sm[O_DOUT + m * 4] = dq;
            }
            {
                float ca1[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, d1[2] = {0.f, 0.f};   // the critic's own gradients from the actor loss are discarded
                const int gnone[1] = {-1};
                ln2_backward<1>(k, c, ep, gnone, b0, slope, ca1, d1);
            }
            stamp(k, 14);
            dh1(k, wQ, wQ, false);
            stamp(k, 15);
            // LayerNorm1 backward of Q1(s, pi) on the rows this CTA owns -> d pi, shared with the whole cluster
            x_begin(k, A * RB);
            for (int q = warp; q < RPW; q += WARPS) {
                float g4[A];
                ln1_rows<CriticL>(k, c, c.critic, 0, 4, b0, slope, q, false, g4);
                const int m = k.crank + 8 * q;
                if (m < RB) {
#pragma unroll
                    for (int j = 0; j < A; ++j) x_post(k, m * A + j, g4[j]);
                }
            }
            {
                const float *xb = x_wait(k);
                if (threadIdx.x < RB * A) {
                    const int m = threadIdx.x >> 2, b = b0 + m;
                    const float g = xb[(m & 7) * XS + threadIdx.x];
                    k.sm[O_AP + threadIdx.x] = g;
                    if (k.crank == 0 && b < B) c.sc.GPI[(size_t)b * A + (threadIdx.x & 3)] = g;
                }
            }
            load_pass(k, c, 2, 0, b0);
            if (hirl) load_pass(k, c, 3, 1, b0);
            __syncthreads();
            stamp(k, 16);
            actor_backward(k, c, st, b0, true, hirl, aa);
            stamp(k, 17);
        }
        if (k.cid < nblk) {
            flush_actor(k, c, aa, true, hirl);
            flush_scalar(k, c, 2, cnt);
            flush_scalar(k, c, 3, sum_q);
        }
    }

    // ===========================================================================================================
    // behaviour cloning (mode 2, BC.py:160-185): expert-state actor pass + its backward
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_bc(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        ActAcc aa;
        aa.cacc[0][0] = aa.cacc[0][1] = aa.cacc[1][0] = aa.cacc[1][1] = 0.f;
        aa.db3[0] = aa.db3[1] = 0.f; aa.bc_sq = aa.fire_sq = 0.f;
        if (k.cid < nblk) fill_pc<ActorL>(k, N_A, c.actor);
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_actor);
            gather_rows(k, c, st, step, b0, false, true);
            {
                L1D *d = l1_table(k);
                if (threadIdx.x == 0) {
                    d[0] = L1D{c.actor, k.sm + O_XBROW, nullptr, BCW, 0, P_AB, 0};
                    gp_table(k)[0] = GP{c.w2p_actor, c.actor + ActorL::B2, P_AB, 3};
                }
                __syncthreads();
                l1_rows(c, d, 1, k.crank, b0, slope);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            const EP ep[1] = {{N_A, A, 3}};
            run_segment(k, c, gp_table(k), 1, b0, c.w2p_actor);   // the slice is requested again for the backward pass
            stats_round(k, ep, 1);
            head_round(k, ep, 0, 1, slope);
            if (threadIdx.x < RB * A) {
                const int m = threadIdx.x >> 2, b = b0 + m;
                if (b < B) c.sc.PIB[(size_t)b * A + (threadIdx.x & 3)] = tanhf(k.sm[O_OUTS + (3 * RB + m) * 4 + (threadIdx.x & 3)]);
            }
            __syncthreads();
            actor_backward(k, c, st, b0, false, true, aa);
        }
        if (k.cid < nblk) flush_actor(k, c, aa, false, true);
    }

    // ===========================================================================================================
    // phases B / D: weight gradients that need every row, Adam, Polyak
    // ===========================================================================================================
    struct Opt { float neg_step, bc2_sqrt; bool polyak; float tau, omt; };
    // one parameter of the internal arena: Adam (torch single-tensor arithmetic), optional Polyak of its target and
    // refresh of the padded W2 images
    static __device__ __forceinline__ void apply(float g, float *w, float *gw, float *m, float *v, float *tgt, const Opt &o, int e, float *img, float *timg,
                                                 int ie) {
        float wv = w[e], mv = m[e], vv = v[e];
        adam_elem(g, wv, mv, vv, o.neg_step, o.bc2_sqrt);
        w[e] = wv; m[e] = mv; v[e] = vv; gw[e] = g;
        if (img) img[ie] = wv;
        if (o.polyak) {
            const float t = polyak_elem(tgt[e], wv, o.tau, o.omt);
            tgt[e] = t;
            if (timg) timg[ie] = t;
        }
    }
    // N parameters at once: all loads first (the arena pointers may alias as far as the compiler knows, so element-wise
    // apply() serialises one L2 round trip per element), then the arithmetic, then the stores
    template <int N>
    static __device__ __forceinline__ void apply_n(const float (&g)[N], float *__restrict__ w, float *__restrict__ gw, float *__restrict__ m, float *__restrict__ v,
                                                   float *__restrict__ tgt, const Opt &o, const int (&e)[N], float *__restrict__ img, float *__restrict__ timg,
                                                   const int (&ie)[N], unsigned okmask = 0xffffffffu) {
        float wv[N], mv[N], vv[N], tv[N];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const bool ok = (okmask >> j) & 1u;
            wv[j] = ok ? __ldcg(w + e[j]) : 0.f; mv[j] = ok ? __ldcg(m + e[j]) : 0.f; vv[j] = ok ? __ldcg(v + e[j]) : 1.f;
            tv[j] = (ok && o.polyak) ? __ldcg(tgt + e[j]) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < N; ++j) adam_elem(g[j], wv[j], mv[j], vv[j], o.neg_step, o.bc2_sqrt);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            if (!((okmask >> j) & 1u)) continue;
            w[e[j]] = wv[j]; m[e[j]] = mv[j]; v[e[j]] = vv[j]; gw[e[j]] = g[j];
            if (img) img[ie[j]] = wv[j];
            if (o.polyak) {
                const float t = polyak_elem(tv[j], wv[j], o.tau, o.omt);
                tgt[e[j]] = t;
                if (timg) timg[ie[j]] = t;
            }
        }
    }
    static __device__ __forceinline__ int nclu_of(const Ctl &k, const AgentCtx &c) {
        const int nblk = (c.B + RB - 1) / RB;
        return nblk < k.ncl ? nblk : k.ncl;
    }
    // Description of one trainable network for phases B / D
    struct Net {
        float *w, *gw, *m, *v, *tgt, *img, *timg;
        int ntowers, tsize, k0, w1t, b1, g1, be1, w2;   // arena geometry (floats)
        int slot_a, slot_b, pass_a, pass_b;             // gradient slots / activation passes of the (up to) two branches; slot_a indexes towers
        float wa, wb;                                   // branch weights
        const float *xa, *xb; int xlda, xldb;           // first-layer inputs of the branches
    };
    // First-layer parameter gradients, 8 output columns per CTA job: db1, dg1, dbe1, dW1[k] summed over the batch, then
    // Adam.  Lanes = 8 columns x 4 row subsets (one 32-byte sector per row and array), warps = 8 more row subsets; the
    // rows' first-layer inputs are staged in shared memory with cp.async (whole 128 / 80-byte rows).  Jobs go to the
    // HIGHEST block indices: with fewer dW2 tiles than 3 x CTAs those are the CTAs with one tile less.
    static __device__ __noinline__ void l1_columns(const AgentCtx &c, const Net &nt, const Opt &o, float *sm) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, col = lane & 7, rs = lane >> 3, B = c.B;
        const int nq = nt.k0 + 3, k0 = nt.k0;
        float *xs_s = sm, *part = sm + 256 * 32;   // [256 rows][xld <= 32] | [8 warps][SA + 3][8]
        for (int job = (int)gridDim.x - 1 - (int)blockIdx.x; job < nt.ntowers * (H1 / 8); job += gridDim.x) {
            const int t = job / (H1 / 8), n = (job - t * (H1 / 8)) * 8 + col;
            float acc[SA + 3];
#pragma unroll
            for (int i = 0; i < SA + 3; ++i) acc[i] = 0.f;
#pragma unroll
            for (int br = 0; br < 2; ++br) {
                const int slot = br ? nt.slot_b : (nt.slot_a < 0 ? -1 : nt.slot_a + t);
                if (slot < 0) continue;
                const float wt = br ? nt.wb : nt.wa;
                const float *xs = br ? nt.xb : nt.xa;
                const int xld = br ? nt.xldb : nt.xlda, xch = xld >> 2;   // 16-byte chunks per row
                const float *dz1 = c.sc.DZ1 + (size_t)slot * B * H1 + n, *dy1 = c.sc.DY1 + (size_t)slot * B * H1 + n, *xh1 = XH1p(c, slot) + n;
                for (int c0 = 0; c0 < B; c0 += 256) {
                    const int rows = min(256, B - c0);
                    __syncthreads();
                    for (int i = threadIdx.x; i < rows * xch; i += CTHREADS) cp16(xs_s + 4 * i, xs + (size_t)c0 * xld + 4 * i, true);
                    cp_commit();
                    float dz[8], dy[8], xh[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {   // all row loads in flight together
                        const int b = 4 * warp + rs + 32 * u;
                        const size_t ro = (size_t)(c0 + min(b, rows - 1)) * H1;
                        const bool ok = b < rows;
                        dz[u] = ok ? wt * __ldcg(dz1 + ro) : 0.f; dy[u] = ok ? wt * __ldcg(dy1 + ro) : 0.f; xh[u] = ok ? __ldcg(xh1 + ro) : 0.f;
                    }
                    cp_wait_all();
                    __syncthreads();
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int b = min(4 * warp + rs + 32 * u, rows - 1);
                        acc[SA] += dz[u];
                        acc[SA + 1] = fmaf(dy[u], xh[u], acc[SA + 1]);
                        acc[SA + 2] += dy[u];
#pragma unroll
                        for (int kk = 0; kk < SA; ++kk)
                            if (kk < k0) acc[kk] = fmaf(dz[u], xs_s[b * xld + kk], acc[kk]);
                    }
                }
            }
            __syncthreads();
#pragma unroll
            for (int i = 0; i < SA + 3; ++i) {
                float v = acc[i];
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (rs == 0) part[(warp * (SA + 3) + i) * 8 + col] = v;
            }
            __syncthreads();
            if ((int)threadIdx.x < nq * 8) {
                const int qi = threadIdx.x >> 3, cc = threadIdx.x & 7, i = qi < k0 ? qi : SA + (qi - k0), nn = n - col + cc;
                float g = 0.f;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) g += part[(w * (SA + 3) + i) * 8 + cc];
                const int off = qi < k0 ? nt.w1t + qi * H1 + nn : qi == k0 ? nt.b1 + nn : qi == k0 + 1 ? nt.g1 + nn : nt.be1 + nn;
                apply(g, nt.w, nt.gw, nt.m, nt.v, nt.tgt, o, t * nt.tsize + off, nullptr, nullptr, 0);
            }
        }
        __syncthreads();
    }
    // dW2 tiles: 64 output features x TK hidden-1 columns, reduction over the batch (per tower up to two row sets =
    // branches, combined with the branch weights), Adam on the finished tile with lanes along k (contiguous rows of W2).
    // A unit = (tile, branch, chunk of <= 256 rows).  The dz2 operand of a unit is one contiguous block of the slice-major
    // scratch (one bulk copy), the h1 operand is staged with cp.async; the NEXT unit's operands are requested as soon as
    // the current unit has been accumulated, so they land while the tile is reduced and Adam runs.
    template <int TK>
    static __device__ __noinline__ unsigned dw2_tiles(const AgentCtx &c, const Net &nt, const Opt &o, float *sm, uint64_t *bar, unsigned phase) {
        constexpr int KL = TK / 4, NKT = H1 / TK;
        float *Dz = sm, *Hs = sm + 256 * NS, *red = Hs + 256 * KS;
        const int B = c.B, tid = threadIdx.x, ntiles = nt.ntowers * CL * NKT;
        const int br0 = nt.slot_a >= 0 ? 0 : 1, br_last = nt.slot_b >= 0 ? 1 : 0;
        auto request = [&](int tile, int br, int c0) {
            const int t = tile / (CL * NKT), rem = tile - t * CL * NKT, ns = rem / NKT, k0 = (rem - ns * NKT) * TK;
            const int slot = br ? nt.slot_b : nt.slot_a + t, pass = br ? nt.pass_b : nt.pass_a + t;
            const int rows = min(256, B - c0), rp = (rows + 7) & ~7;
            const float *dzs = c.sc.DZ2S + ((size_t)(slot * CL + ns) * c.Bpad + c0) * NS;
            const float *h1p = c.sc.H1P + (size_t)pass * c.Bpad * WLD;
            if (tid == 0) {
                mbar_expect_tx(bar, (uint32_t)rows * NS * 4u);
                bulk_g2s(Dz, dzs, (uint32_t)rows * NS * 4u, bar);
            }
            for (int i = rows * NS + tid; i < rp * NS; i += CTHREADS) Dz[i] = 0.f;   // rows past the batch
            constexpr int CH = TK / 4;
            for (int cidx = tid; cidx < rp * CH; cidx += CTHREADS) {
                const int r = cidx / CH, q = cidx - r * CH, b = c0 + r;
                cp16(Hs + r * TK + 4 * q, h1p + (size_t)min(b, B - 1) * WLD + k0 + 4 * q, b < B);
            }
            cp_commit();
        };
        int tile = blockIdx.x, br = br0, c0 = 0;
        if (tile < ntiles) request(tile, br, c0);
        float g[KL], acc[8][KL];
        while (tile < ntiles) {
            if (br == br0 && c0 == 0) {
#pragma unroll
                for (int j = 0; j < KL; ++j) g[j] = 0.f;
            }
            if (c0 == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < KL; ++j) acc[i][j] = 0.f;
            }
            const int rows = min(256, B - c0), rp = (rows + 7) & ~7;
            mbar_wait(bar, phase & 1u, 7);
            phase += 1u;
            cp_wait_all();
            __syncthreads();
            dw_accumulate<TK>(Dz, Hs, rp, acc);
            // next unit
            int ntile = tile, nbr = br, nc0 = c0 + 256;
            if (nc0 >= B) { nc0 = 0; nbr = br + 1; if (nbr > br_last) { nbr = br0; ntile = tile + gridDim.x; } }
            __syncthreads();   // every warp is done with the staged operands
            if (ntile < ntiles) request(ntile, nbr, nc0);
            if (nc0 == 0) {   // this branch of the tile is complete
                float gb[KL];
                dw_reduce<TK>(acc, red, gb);
                const float wt = br ? nt.wb : nt.wa;
#pragma unroll
                for (int j = 0; j < KL; ++j) g[j] = fmaf(wt, gb[j], g[j]);
                if (br == br_last) {
                    const int t = tile / (CL * NKT), rem = tile - t * CL * NKT, ns = rem / NKT, k0 = (rem - ns * NKT) * TK;
                    const int kk = k0 + tid % TK, nn = NS * ns + tid / TK;
                    int e[KL], ie[KL];
#pragma unroll
                    for (int j = 0; j < KL; ++j) {
                        const int n = nn + (CTHREADS / TK) * j;
                        e[j] = t * nt.tsize + nt.w2 + n * H1 + kk;
                        ie[j] = (t * H2 + n) * WLD + kk;
                    }
                    apply_n<KL>(g, nt.w, nt.gw, nt.m, nt.v, nt.tgt, o, e, nt.img, nt.timg, ie);
                }
            }
            tile = ntile; br = nbr; c0 = nc0;
        }
        return phase;
    }

    static __device__ __forceinline__ void phase_b(Ctl &k, const AgentCtx &c, const hirl_step_t &st, bool tk16) {
        const int nclu = nclu_of(k, c);
        const Opt o{st.critic_neg_step, st.critic_bc2_sqrt, false, 0.f, 0.f};
        // small second-layer parameters: fixed-order sum of the per-cluster partials, then Adam
        for (int e = blockIdx.x * CTHREADS + threadIdx.x; e < 2 * 4 * H2 + 2; e += gridDim.x * CTHREADS) {
            float pv[MAX_CL];
#pragma unroll
            for (int cc = 0; cc < MAX_CL; ++cc) pv[cc] = cc < nclu ? __ldcg(c.sc.PART2C + (size_t)cc * PC_WORDS + e) : 0.f;   // all in flight together
            float g = 0.f;
#pragma unroll
            for (int cc = 0; cc < MAX_CL; ++cc) g += pv[cc];
            int off;
            if (e < 2 * 4 * H2) {
                const int t = e / (4 * H2), q = (e / H2) & 3, n = e & (H2 - 1);
                off = t * CriticL::SIZE + (q == 0 ? CriticL::B2 : q == 1 ? CriticL::G2 : q == 2 ? CriticL::BE2 : CriticL::W3) + n;
            } else {
                off = (e - 2 * 4 * H2) * CriticL::SIZE + CriticL::B3;
            }
            apply(g, c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, o, off, nullptr, nullptr, 0);
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {   // critic_loss = mse(Q1,y) + mse(Q2,y)  (HIRL.py:284)
            float s0 = 0.f, s1 = 0.f;
            for (int cc = 0; cc < nclu; ++cc) { s0 += __ldcg(c.sc.LP + cc * LPW); s1 += __ldcg(c.sc.LP + cc * LPW + 1); }
            if (threadIdx.x == 0) c.carry[0] = __fadd_rn(s0 / (float)c.B, s1 / (float)c.B);
        }
        stamp(k, 10);
        const Net nt{c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, c.w2p_critic, nullptr,
                     2, CriticL::SIZE, SA, CriticL::W1T, CriticL::B1, CriticL::G1, CriticL::BE1, CriticL::W2,
                     0, -1, P_C1, 0, 1.f, 0.f, c.sc.X, nullptr, ROW, 0};
        l1_columns(c, nt, o, k.sm);
        stamp(k, 11);
        k.tph = tk16 ? dw2_tiles<16>(c, nt, o, k.sm, k.xbar + 2, k.tph) : dw2_tiles<32>(c, nt, o, k.sm, k.xbar + 2, k.tph);
    }

    static __device__ __forceinline__ void phase_d(Ctl &k, const AgentCtx &c, const hirl_step_t &st, bool tk16) {
        const int nclu = nclu_of(k, c), B = c.B;
        const bool rl = st.mode != 2, bc = st.mode != 1;
        // the behaviour-cloning weight of this call (HIRL.py:292-308), evaluated identically by every thread
        float count = 0.f;
        double wd = (double)st.bc_weight_now;
        if (st.mode == 2) wd = 1.0;
        else if (st.mode != 0) wd = 0.0;
        else if (st.bc_weight_now == 100.f) {
            for (int cc = 0; cc < nclu; ++cc) count += __ldcg(c.sc.LP + cc * LPW + 2);
            const float mean = st.force_soft_weight >= 0.f ? st.force_soft_weight : count / (float)B;
            wd = (double)mean + (double)st.bc_warm_up;
        }
        if (wd > 1.0) wd = 1.0;
        const float wbc = (float)wd, wrl = (float)(1.0 - wd);
        const Opt o{st.actor_neg_step, st.actor_bc2_sqrt, st.do_polyak != 0, c.tau, (float)(1.0 - (double)c.tau)};
        for (int e = blockIdx.x * CTHREADS + threadIdx.x; e < 7 * H2 + A; e += gridDim.x * CTHREADS) {
            float ga = 0.f, gb = 0.f;
            const int ea = e < 7 * H2 ? e : 2 * 7 * H2 + (e - 7 * H2), eb = e < 7 * H2 ? 7 * H2 + e : 2 * 7 * H2 + 4 + (e - 7 * H2);
            {
                float pa[MAX_CL], pb[MAX_CL];
#pragma unroll
                for (int cc = 0; cc < MAX_CL; ++cc) {
                    pa[cc] = (rl && cc < nclu) ? __ldcg(c.sc.PART2A + (size_t)cc * PA_WORDS + ea) : 0.f;
                    pb[cc] = (bc && cc < nclu) ? __ldcg(c.sc.PART2A + (size_t)cc * PA_WORDS + eb) : 0.f;
                }
#pragma unroll
                for (int cc = 0; cc < MAX_CL; ++cc) { ga += pa[cc]; gb += pb[cc]; }
            }
            const float g = fmaf(wrl, ga, wbc * gb);
            int off;
            if (e < 7 * H2) {
                const int q = e / H2, n = e & (H2 - 1);
                off = (q == 0 ? ActorL::B2 : q == 1 ? ActorL::G2 : q == 2 ? ActorL::BE2 : ActorL::W3 + (q - 3) * H2) + n;
            } else {
                off = ActorL::B3 + (e - 7 * H2);
            }
            apply(g, c.actor, c.g_actor, c.m_actor, c.v_actor, c.tactor, o, off, nullptr, nullptr, 0);
        }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
            float sq = 0.f, sb = 0.f, sf = 0.f;
            for (int cc = 0; cc < nclu; ++cc) { sq += __ldcg(c.sc.LP + cc * LPW + 3); sb += __ldcg(c.sc.LP + cc * LPW + 4); sf += __ldcg(c.sc.LP + cc * LPW + 5); }
            const float rl_loss = st.mode == 2 ? 0.f : -(sq / (float)B);                                  // HIRL.py:297
            float bc_loss = 0.f, fire = 0.f;
            if (st.mode == 0) {
                bc_loss = __fmul_rn(sb / (float)(A * B), c.loss_lambda);                                   // HIRL.py:311
                fire = sf / (float)B;                                                                      // HIRL.py:319 (x lambda on the host)
            } else if (st.mode == 2) {
                bc_loss = sb / (float)(A * B);                                                             // BC.py:179
            }
            c.carry[1] = __fadd_rn(__fmul_rn(bc_loss, wbc), __fmul_rn(rl_loss, wrl));                      // HIRL.py:321
            c.carry[2] = bc_loss; c.carry[3] = rl_loss; c.carry[4] = fire; c.carry[5] = wbc; c.carry[6] = count;
        }
        stamp(k, 20);
        const Net nt{c.actor, c.g_actor, c.m_actor, c.v_actor, c.tactor, c.w2p_actor, c.w2p_tactor,
                     1, ActorL::SIZE, S, ActorL::W1T, ActorL::B1, ActorL::G1, ActorL::BE1, ActorL::W2,
                     rl ? 2 : -1, bc ? 3 : -1, P_A, P_AB, wrl, wbc, c.sc.X, c.sc.XB, ROW, BCW};
        l1_columns(c, nt, o, k.sm);
        if (o.polyak && st.mode != 2) {   // target critic <- critic (HIRL.py:328-329); the critic is final since phase B
            for (int e = blockIdx.x * CTHREADS + threadIdx.x; e < CRITIC_SIZE; e += gridDim.x * CTHREADS) {
                const float tv = polyak_elem(c.tcritic[e], c.critic[e], o.tau, o.omt);
                c.tcritic[e] = tv;
                const int t = e / CriticL::SIZE, off = e - t * CriticL::SIZE;
                if (off >= CriticL::W2 && off < CriticL::B2) {
                    const int i = off - CriticL::W2, n = i / H1, kk = i - n * H1;
                    c.w2p_tcritic[(size_t)(t * H2 + n) * WLD + kk] = tv;
                }
            }
        }
        stamp(k, 21);
        k.tph = tk16 ? dw2_tiles<16>(c, nt, o, k.sm, k.xbar + 2, k.tph) : dw2_tiles<32>(c, nt, o, k.sm, k.xbar + 2, k.tph);
    }
};

// ---------------------------------------------------------------------------------------------------------------
template <int RI>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CTHREADS, 1) hirl_cluster_kernel(const __grid_constant__ ClusterLaunch L) {
    using K = Ker<RI>;
    extern __shared__ float4 smem4[];
    __shared__ AgentCtx ctx;
    __shared__ hirl_step_t st;
    __shared__ uint64_t bars[7];
    {
        const int *src = reinterpret_cast<const int *>(&L.agent);
        int *dst = reinterpret_cast<int *>(&ctx);
        for (int i = threadIdx.x; i < (int)(sizeof(AgentCtx) / sizeof(int)); i += CTHREADS) dst[i] = src[i];
    }
    typename K::Ctl k;
    k.sm = reinterpret_cast<float *>(smem4);
    k.wbar = bars; k.abar = bars + 2; k.xbar = bars + 4;
    k.crank = (int)cluster_ctarank(); k.cid = (int)cluster_id_x(); k.ncl = L.ncl;
    k.wi = k.wo = k.ai = k.ao = 0u; k.xr = 0u; k.tph = 0u;
    k.flags = L.bar_flags;
    k.gen = ld_relaxed_u32(k.flags + k.cid * FLAG_STRIDE);
    k.pr = nullptr;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 7; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();   // peers' barriers are initialised before anybody posts to them
    const bool tk16 = (int)gridDim.x < 2 * CL * (H1 / 32);
    for (int step = 0; step < L.k_steps; ++step) {
        __syncthreads();
        if (threadIdx.x < (int)(sizeof(hirl_step_t) / sizeof(int)))
            reinterpret_cast<int *>(&st)[threadIdx.x] = reinterpret_cast<const int *>(ctx.steps + step)[threadIdx.x];
        __syncthreads();
        k.pr = (L.prof && step < PROF_STEPS && blockIdx.x == 0 && threadIdx.x == 0) ? L.prof + (size_t)step * PROF_WORDS : nullptr;
        K::stamp(k, 0);
        long long *gpr = (L.prof && step < PROF_STEPS) ? L.prof + (size_t)PROF_STEPS * PROF_WORDS + (size_t)step * 4 * MAX_CL : nullptr;
        const bool bc_only = st.mode == 2;
#define HIRL_PH(i, call)       \
    call;                      \
    K::gstamp(k, gpr, i);      \
    K::stamp(k, 24 + 2 * (i)); \
    K::grid_sync(k);           \
    K::stamp(k, 25 + 2 * (i));
        if (!bc_only) {
            HIRL_PH(0, K::phase_a(k, ctx, st, step))
            HIRL_PH(1, K::phase_b(k, ctx, st, tk16))
            if (st.do_actor) {
                HIRL_PH(2, K::phase_c(k, ctx, st, step))
                HIRL_PH(3, K::phase_d(k, ctx, st, tk16))
            }
        } else {
            HIRL_PH(2, K::phase_bc(k, ctx, st, step))
            HIRL_PH(3, K::phase_d(k, ctx, st, tk16))
        }
#undef HIRL_PH
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x < HIRL_OUT_SCALARS) ctx.out_scalars[(size_t)step * HIRL_OUT_SCALARS + threadIdx.x] = ctx.carry[threadIdx.x];
    }
}

}  // namespace cl
}  // namespace hirl
