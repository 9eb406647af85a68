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
// image of the matrix) is streamed into shared memory while the previous pass computes.  FP32 FMA throughout.
// Everything is reduced in a fixed order: results are run-to-run deterministic.
#pragma once
#include "cluster_gemm.cuh"
#include "update.cuh"

namespace hirl {
namespace cl {

constexpr int MAX_CL = 24;                 // clusters per launch (B200: 15 or 16 co-resident clusters of 8)
constexpr int PC_WORDS = 2 * 4 * H2 + 32;  // per-cluster partial sums of the critic's small second-layer parameters
constexpr int PA_WORDS = 2 * 7 * H2 + 32;  // same for the actor, per branch
constexpr int LPW = 16;                    // per-cluster loss / count words

struct ClusterLaunch {
    AgentCtx agent;
    unsigned *bar_flags;   // [MAX_CL][32] cluster flags of the grid barrier (generations keep counting across launches)
    int k_steps, ncl;
    long long *prof;       // [PROF_STEPS][8][2] SM-clock stamps of CTA 0 (phase end / barrier end), may be null
};

__device__ __forceinline__ float sum8(float v) {   // lanes l, l^1, l^2, l^4: fixed tree, identical on every lane
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
template <int RI>
struct Ker {
    static constexpr int RB = 2 * RI, RPW = (RB + 7) / 8, RIP = (RI + 3) & ~3, XS = 12 * RB;
    // shared-memory carve-up (floats)
    static constexpr int O_W0 = 0, O_W1 = W_FLOATS, O_A0 = 2 * W_FLOATS, O_A1 = O_A0 + RB * WLD, O_RED = O_A1 + RB * WLD,
                         O_DZS = O_RED + 4 * RB * RLD2, O_ROWS = O_DZS + 2 * NS * 2 * RIP, O_XB = O_ROWS + 2 * 3 * H1, O_XROW = O_XB + 2 * CL * XS,
                         O_XBROW = O_XROW + RB * 32, O_AP = O_XBROW + RB * 20, O_MISC = O_AP + RB * 4, TOTAL_A = O_MISC + 64;
    static constexpr int TOTAL_B = 256 * NS + 256 * KS + 4 * KS * NS;   // phase B/D: Dz, Hs, reduction
    static constexpr int TOTAL = TOTAL_A > TOTAL_B ? TOTAL_A : TOTAL_B;
    static constexpr size_t SMEM_BYTES = (size_t)TOTAL * 4;
    static_assert(4 * RB * RLD2 >= RB * H1 && 2 * RB * WLD >= RB * H1, "dH1 partial buffers alias the reduction / activation buffers");
    static_assert(4 * RB * RLD2 >= 8 * 8 * NS, "column-sum staging aliases the reduction buffer");

    // -----------------------------------------------------------------------------------------------------------
    struct Ctl {
        float *sm;
        uint64_t *wbar, *abar, *xbar;
        int crank, cid, ncl;
        unsigned wi, wo, ai, ao;   // weight-slice / activation-block loads issued and consumed (buffer = index & 1)
        unsigned xr;               // exchange rounds done (parity = xr & 1)
        unsigned gen;
        unsigned *flags;
    };

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
    // forward GEMM of the next activation block against the current weight slice; sums land in the reduction buffer
    static __device__ __forceinline__ void gemm(Ctl &k, bool last_use_of_w) {
        mbar_wait(k.wbar + (k.wo & 1u), (k.wo >> 1) & 1u, 1);
        mbar_wait(k.abar + (k.ao & 1u), (k.ao >> 1) & 1u, 2);
        fwd_pass<RI>(k.sm + O_A0 + (k.ao & 1u) * RB * WLD, k.sm + (k.wo & 1u) * W_FLOATS, k.sm + O_RED);
        k.ao += 1u;
        if (last_use_of_w) k.wo += 1u;
    }

    // ---- DSMEM exchange rounds ---------------------------------------------------------------------------------
    static __device__ __forceinline__ void x_begin(Ctl &k, int values_expected) {
        __syncthreads();   // everybody has consumed the buffer of two rounds ago
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

    // ---- per-row helpers in the epilogue mapping: warp w holds rows m = w + 8 r, lane l columns l, l + 32 of the slice
    struct Head { const float *b2, *g2, *be2, *w3, *b3; };
    template <class L>
    static __device__ __forceinline__ Head head_of(const float *tw) { return Head{tw + L::B2, tw + L::G2, tw + L::BE2, tw + L::W3, tw + L::B3}; }

    struct Pass {           // forward state of one pass, this thread's RPW x 2 elements
        float z[RPW][2];    // pre-LayerNorm (bias included), later xhat
        float rstd[RPW];
    };
    static __device__ __forceinline__ void read_z(const Ctl &k, const Head &h, Pass &p) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RPW; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c)
                p.z[r][c] = (warp + 8 * r < RB) ? fwd_value<RI>(k.sm + O_RED, warp + 8 * r, lane + 32 * c) + h.b2[NS * k.crank + lane + 32 * c] : 0.f;
    }
    // posts (mean_i, M2_i) of the CTA's 64 columns for every row of pass index pi
    static __device__ __forceinline__ void post_stats(const Ctl &k, int pi, const Pass &p) {
        const int warp = threadIdx.x >> 5;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
            const float mean = warp_sum(p.z[r][0] + p.z[r][1]) * (1.f / NS);
            const float d0 = p.z[r][0] - mean, d1 = p.z[r][1] - mean;
            const float m2 = warp_sum(fmaf(d0, d0, d1 * d1));
            x_post(k, (pi * RB + m) * 2, mean);
            x_post(k, (pi * RB + m) * 2 + 1, m2);
        }
    }
    // combines the 8 CTAs' statistics (Chan's parallel variance), normalises: z -> xhat, fills rstd
    static __device__ __forceinline__ void finish_stats(const Ctl &k, const float *xb, int pi, Pass &p) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
            const float mi = xb[(lane & 7) * XS + (pi * RB + m) * 2], m2i = xb[(lane & 7) * XS + (pi * RB + m) * 2 + 1];
            const float mean = sum8(mi) * (1.f / CL);
            const float d = mi - mean;
            const float m2 = sum8(fmaf((float)NS * d, d, m2i));
            const float rstd = 1.f / sqrtf(m2 * (1.f / H2) + LN_EPS);
            p.rstd[r] = rstd;
            p.z[r][0] = (p.z[r][0] - mean) * rstd;
            p.z[r][1] = (p.z[r][1] - mean) * rstd;
        }
    }
    // y = LayerNorm output, h = activation; posts the partial head dot products (slot base given by the caller)
    template <int NOUT>
    static __device__ __forceinline__ void post_head(const Ctl &k, const Head &h, float slope, int slot0, const Pass &p) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n0 = NS * k.crank + lane;
        const float g0 = h.g2[n0], g1 = h.g2[n0 + 32], e0 = h.be2[n0], e1 = h.be2[n0 + 32];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
            const float h0 = act_f(fmaf(p.z[r][0], g0, e0), slope), h1 = act_f(fmaf(p.z[r][1], g1, e1), slope);
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                const float d = warp_sum(fmaf(h0, h.w3[o * H2 + n0], h1 * h.w3[o * H2 + n0 + 32]));
                x_post(k, slot0 + m * NOUT + o, d);
            }
        }
    }
    template <int NOUT>
    static __device__ __forceinline__ float get_head(const float *xb, const Head &h, int slot0, int m, int o) {
        const int lane = threadIdx.x & 31;
        return sum8(xb[(lane & 7) * XS + slot0 + m * NOUT + o]) + h.b3[o];
    }

    // backward through head + LayerNorm2 of one pass, first half: dy, dxhat and the row sums to exchange
    struct Bwd { float dy[RPW][2], dxh[RPW][2]; };
    template <int NOUT>
    static __device__ __forceinline__ void ln2_bwd_post(const Ctl &k, const Head &h, float slope, int slot0, const Pass &p, const float (&dout)[RPW][4],
                                                        Bwd &bw) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n0 = NS * k.crank + lane;
        const float g0 = h.g2[n0], g1 = h.g2[n0 + 32], e0 = h.be2[n0], e1 = h.be2[n0 + 32];
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
            float dy0 = 0.f, dy1 = 0.f;
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                dy0 = fmaf(dout[r][o], h.w3[o * H2 + n0], dy0);
                dy1 = fmaf(dout[r][o], h.w3[o * H2 + n0 + 32], dy1);
            }
            dy0 *= act_g(fmaf(p.z[r][0], g0, e0), slope);
            dy1 *= act_g(fmaf(p.z[r][1], g1, e1), slope);
            bw.dy[r][0] = dy0; bw.dy[r][1] = dy1;
            bw.dxh[r][0] = dy0 * g0; bw.dxh[r][1] = dy1 * g1;
            x_post(k, slot0 + 2 * m, warp_sum(bw.dxh[r][0] + bw.dxh[r][1]));
            x_post(k, slot0 + 2 * m + 1, warp_sum(fmaf(bw.dxh[r][0], p.z[r][0], bw.dxh[r][1] * p.z[r][1])));
        }
    }
    // second half: dz = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)); written into bw.dxh
    static __device__ __forceinline__ void ln2_bwd_finish(const float *xb, int slot0, const Pass &p, Bwd &bw) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
            const float s1 = sum8(xb[(lane & 7) * XS + slot0 + 2 * m]) * (1.f / H2), s2 = sum8(xb[(lane & 7) * XS + slot0 + 2 * m + 1]) * (1.f / H2);
            bw.dxh[r][0] = p.rstd[r] * (bw.dxh[r][0] - s1 - p.z[r][0] * s2);
            bw.dxh[r][1] = p.rstd[r] * (bw.dxh[r][1] - s1 - p.z[r][1] * s2);
        }
    }
    // dz2 of this CTA's slice: shared copy for the dH1 pass ([n][mg][RIP]) and the slice-major global scratch for dW2
    static __device__ __forceinline__ void store_dz(const Ctl &k, const AgentCtx &c, int which, int slot, int b0, const Bwd &bw) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float *dzs = k.sm + O_DZS + which * NS * 2 * RIP;
        float *g = c.sc.DZ2S + ((size_t)(slot * CL + k.crank) * c.Bpad) * NS;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const int m = warp + 8 * r;
            if (m >= RB) continue;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int n = lane + 32 * cc;
                dzs[(n * 2 + (m & 1)) * RIP + (m >> 1)] = bw.dxh[r][cc];
                if (b0 + m < c.B) g[(size_t)(b0 + m) * NS + n] = bw.dxh[r][cc];
            }
        }
    }
    // stages this warp's row-summed column values: cs[warp][q][64]
    static __device__ __forceinline__ void stage_col(const Ctl &k, int q, int nq, const float (&v)[RPW][2]) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int r = 0; r < RPW; ++r)
            if (warp + 8 * r < RB) { s0 += v[r][0]; s1 += v[r][1]; }
        float *cs = k.sm + O_RED + (warp * nq + q) * NS;
        cs[lane] = s0;
        cs[lane + 32] = s1;
    }
    static __device__ __forceinline__ float take_col(const Ctl &k, int item, int nq) {   // item = q * 64 + col
        const float *cs = k.sm + O_RED + item;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) s += cs[w * nq * NS];
        return s;
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

    // one first-layer row: x from shared memory, h1 into the padded activation rows (csrc/update.cuh: l1_row)
    template <class L>
    static __device__ __forceinline__ void l1_inst(const AgentCtx &c, const float *tw, const float *xsrc, int pass, int b, float slope) {
        const int lane = threadIdx.x & 31, slot = bwd_slot(pass);
        float x[L::k0];
#pragma unroll
        for (int kk = 0; kk < L::k0; ++kk) x[kk] = xsrc[kk];
        const bool keep = slot >= 0 && b < c.B;
        l1_row<L>(tw, x, slope, lane, c.sc.H1P + ((size_t)pass * c.Bpad + b) * WLD, keep ? XH1p(c, slot) + (size_t)b * H1 : nullptr,
                  keep ? RS1p(c, slot) + b : nullptr);
    }

    // LayerNorm1 backward of the rows this CTA owns (m = crank + 8 q); dh1 rows come from the reduce-scatter buffer.
    // Writes dz1 / dy1 for the first-layer gradient pass; optionally d(action) (HIRL.py:296: Q1's gradient w.r.t. pi).
    template <class L>
    static __device__ __forceinline__ void ln1_rows(const Ctl &k, const AgentCtx &c, const float *tw, int which, int slot, int b0, float slope, int q,
                                                    float *gpi4, bool store = true) {
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
        if (store) {   // (Q1(s,pi)'s own first-layer gradient is discarded: only d pi is needed)
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

    // dH1 of two passes (partials -> reduce-scatter) ; results in rows[which][q][256]
    static __device__ __forceinline__ void dh1_two(Ctl &k, const float *w0, const float *w1, bool two) {
        __syncthreads();   // dzs complete, column-sum staging consumed
        bwd_pass<RI>(k.sm + O_DZS, w0, k.sm + O_RED);
        if (two) bwd_pass<RI>(k.sm + O_DZS + NS * 2 * RIP, w1, k.sm + O_A0);
        cluster_sync_all();
        reduce_scatter_rows<RB>(k.sm + O_RED, k.crank, k.sm + O_ROWS);
        if (two) reduce_scatter_rows<RB>(k.sm + O_A0, k.crank, k.sm + O_ROWS + 3 * H1);
        __syncthreads();
        // the partial buffers are overwritten only after a later cluster-wide barrier (next first-layer exchange or
        // the grid barrier), i.e. after every peer has finished these remote reads
    }

    // ---- actor backward for one row block (HIRL.py:295-321; BC.py:177-183), shared by phase C and the BC mode -----
    // branch 0: RL branch on the batch states (pass A), upstream d pi from the critic (unit scale, the (1-w) factor is
    // applied in phase D); branch 1: cloning branch on the expert states (pass AB).
    struct ColAcc { float a[2][2]; float db3[2]; float bc_sq, fire_sq; };
    static __device__ __forceinline__ void actor_backward(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int b0, bool rl, bool bc, const float (&gpi)[RPW][4],
                                                          const float *w_actor, ColAcc &ca) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B;
        const float slope = slope_of(c);
        const Head h = head_of<ActorL>(c.actor);
        const float bc_scale = (2.f / (float)(A * B)) * (st.mode == 2 ? 1.f : c.loss_lambda);
        Pass p[2];
        Bwd bw[2];
        float dpre[2][RPW][4];
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            if (!(br ? bc : rl)) continue;
            const float *pis = br ? c.sc.PIB : c.sc.PI;
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const int m = warp + 8 * r, b = b0 + m;
                const bool live = m < RB && b < B;
                const size_t row = (size_t)br * c.Bpad + (live ? b : 0);
                p[br].z[r][0] = live ? c.sc.XH2[row * H2 + NS * k.crank + lane] : 0.f;
                p[br].z[r][1] = live ? c.sc.XH2[row * H2 + NS * k.crank + lane + 32] : 0.f;
                p[br].rstd[r] = live ? c.sc.RS2[row] : 0.f;
#pragma unroll
                for (int j = 0; j < A; ++j) {
                    const float pi = live ? pis[(size_t)b * A + j] : 0.f;
                    float dpi;
                    if (br == 0) {
                        dpi = gpi[r][j];
                    } else {
                        const float diff = pi - k.sm[O_XBROW + (m < RB ? m : 0) * BCW + S + j];
                        if (live) {
                            ca.bc_sq = fmaf(diff, diff, ca.bc_sq);
                            if (j == 3) ca.fire_sq = fmaf(diff, diff, ca.fire_sq);
                        }
                        dpi = bc_scale * diff;
                    }
                    dpre[br][r][j] = live ? dpi * (1.f - pi * pi) : 0.f;   // tanh backward
                }
            }
        }
        x_begin(k, ((rl ? 1 : 0) + (bc ? 1 : 0)) * 2 * RB * CL);
        if (rl) ln2_bwd_post<A>(k, h, slope, 0, p[0], dpre[0], bw[0]);
        if (bc) ln2_bwd_post<A>(k, h, slope, 2 * RB, p[1], dpre[1], bw[1]);
        const float *xb = x_wait(k);
        if (rl) ln2_bwd_finish(xb, 0, p[0], bw[0]);
        if (bc) ln2_bwd_finish(xb, 2 * RB, p[1], bw[1]);
        // column sums of the small second-layer parameters, one branch at a time: db2, dg2, dbe2, dW3[4]
        const int n0 = NS * k.crank + lane;
        const float g0 = h.g2[n0], g1 = h.g2[n0 + 32], e0 = h.be2[n0], e1 = h.be2[n0 + 32];
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            if (!(br ? bc : rl)) continue;
            __syncthreads();
            float t[RPW][2];
            stage_col(k, 0, 7, bw[br].dxh);
#pragma unroll
            for (int r = 0; r < RPW; ++r) { t[r][0] = bw[br].dy[r][0] * p[br].z[r][0]; t[r][1] = bw[br].dy[r][1] * p[br].z[r][1]; }
            stage_col(k, 1, 7, t);
            stage_col(k, 2, 7, bw[br].dy);
#pragma unroll
            for (int j = 0; j < A; ++j) {
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    t[r][0] = dpre[br][r][j] * act_f(fmaf(p[br].z[r][0], g0, e0), slope);
                    t[r][1] = dpre[br][r][j] * act_f(fmaf(p[br].z[r][1], g1, e1), slope);
                }
                stage_col(k, 3 + j, 7, t);
            }
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < A; ++j) {
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < RPW; ++r) s += dpre[br][r][j];
                    k.sm[O_MISC + warp * 4 + j] = s;
                }
            }
            __syncthreads();
            ca.a[br][0] += take_col(k, threadIdx.x, 7);
            if (threadIdx.x < 7 * NS - CTHREADS) ca.a[br][1] += take_col(k, threadIdx.x + CTHREADS, 7);
            if (threadIdx.x < A) {
                float s = 0.f;
                for (int w = 0; w < WARPS; ++w) s += k.sm[O_MISC + w * 4 + threadIdx.x];
                ca.db3[br] += s;
            }
        }
        __syncthreads();
        if (rl) store_dz(k, c, rl && bc ? 0 : 0, 2, b0, bw[0]);
        if (bc) store_dz(k, c, rl ? 1 : 0, 3, b0, bw[1]);
        // dH1 of both branches against the actor's weight slice, LayerNorm1 backward of the rows this CTA owns
        mbar_wait(k.wbar + (k.wo & 1u), (k.wo >> 1) & 1u, 4);
        dh1_two(k, w_actor, w_actor, rl && bc);
        k.wo += 1u;
        const int nb = (rl ? 1 : 0) + (bc ? 1 : 0);
        for (int inst = warp; inst < nb * RPW; inst += WARPS) {
            const int which = inst / RPW, q = inst - which * RPW;
            const int slot = (which == 0 && rl) ? 2 : 3;
            ln1_rows<ActorL>(k, c, c.actor, which, slot, b0, slope, q, nullptr);
        }
    }
    static __device__ __forceinline__ void flush_actor_cols(const Ctl &k, const AgentCtx &c, const ColAcc &ca, bool rl, bool bc) {
        float *part = c.sc.PART2A + (size_t)k.cid * PA_WORDS;
#pragma unroll
        for (int br = 0; br < 2; ++br) {
            if (!(br ? bc : rl)) continue;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int item = threadIdx.x + u * CTHREADS;
                if (item < 7 * NS) part[(br * 7 + item / NS) * H2 + NS * k.crank + (item & (NS - 1))] = u ? ca.a[br][1] : ca.a[br][0];
            }
            if (k.crank == 0 && threadIdx.x < A) part[2 * 7 * H2 + br * 4 + threadIdx.x] = ca.db3[br];
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

    // ===========================================================================================================
    // phase A (modes 0 / 1): everything of a call that is local to a row block, up to the critic's dz2 / dz1
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_a(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        const bool act = st.do_actor != 0, hirl = st.mode == 0;
        const Head hA = head_of<ActorL>(c.actor), hTA = head_of<ActorL>(c.tactor), hBC = head_of<ActorL>(c.bc);
        const Head hC[2] = {head_of<CriticL>(c.critic), head_of<CriticL>(c.critic + CriticL::SIZE)};
        const Head hT[2] = {head_of<CriticL>(c.tcritic), head_of<CriticL>(c.tcritic + CriticL::SIZE)};
        float col[2] = {0.f, 0.f}, db3[2] = {0.f, 0.f}, loss[2] = {0.f, 0.f};
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_tactor);
            gather_rows(k, c, st, step, b0, true, act && hirl);
            __syncthreads();
            // ---- first layers of the passes whose inputs are known (rows m = crank + 8 q of the block)
            {
                const int np = act ? (hirl ? 6 : 4) : 3;
                for (int inst = warp; inst < np * RPW; inst += WARPS) {
                    const int pi = inst / RPW, q = inst - pi * RPW, m = k.crank + 8 * q, b = b0 + m;
                    if (m >= RB) continue;
                    const float *xr = k.sm + O_XROW + m * 32;
                    if (pi == 0) l1_inst<ActorL>(c, c.tactor, xr + SA, P_TA, b, slope);
                    else if (pi == 1) l1_inst<CriticL>(c, c.critic, xr, P_C1, b, slope);
                    else if (pi == 2) l1_inst<CriticL>(c, c.critic + CriticL::SIZE, xr, P_C2, b, slope);
                    else if (pi == 3) l1_inst<ActorL>(c, c.actor, xr, P_A, b, slope);
                    else if (pi == 4) l1_inst<ActorL>(c, c.actor, k.sm + O_XBROW + m * BCW, P_AB, b, slope);
                    else l1_inst<ActorL>(c, c.bc, xr, P_BC, b, slope);
                }
            }
            fence_proxy_async_all();
            cluster_sync_all();
            const size_t rowoff = (size_t)b0 * WLD, pstride = (size_t)c.Bpad * WLD;
            // ---- segment 1: TA [A, AB, BC]
            Pass pTA, pA, pAB, pBC;
            lda(k, c.sc.H1P + P_TA * pstride + rowoff);
            if (act) { ldw(k, c.w2p_actor); lda(k, c.sc.H1P + P_A * pstride + rowoff); }
            else ldw(k, c.w2p_tcritic);
            gemm(k, true);
            read_z(k, hTA, pTA);
            if (act) {
                if (hirl) { ldw(k, c.w2p_bc); lda(k, c.sc.H1P + P_AB * pstride + rowoff); }
                else ldw(k, c.w2p_tcritic);
                gemm(k, !hirl);
                read_z(k, hA, pA);
                if (hirl) {
                    lda(k, c.sc.H1P + P_BC * pstride + rowoff);
                    gemm(k, true);
                    read_z(k, hA, pAB);
                    ldw(k, c.w2p_tcritic);
                    gemm(k, true);
                    read_z(k, hBC, pBC);
                }
            }
            ldw(k, c.w2p_tcritic + (size_t)H2 * WLD);
            // ---- LayerNorm2 + heads of segment 1
            const int np1 = act ? (hirl ? 4 : 2) : 1;
            x_begin(k, np1 * 2 * RB * CL);
            post_stats(k, 0, pTA);
            if (act) post_stats(k, 1, pA);
            if (act && hirl) { post_stats(k, 2, pAB); post_stats(k, 3, pBC); }
            {
                const float *xb = x_wait(k);
                finish_stats(k, xb, 0, pTA);
                if (act) finish_stats(k, xb, 1, pA);
                if (act && hirl) { finish_stats(k, xb, 2, pAB); finish_stats(k, xb, 3, pBC); }
            }
            x_begin(k, (act ? 2 : 1) * A * RB * CL);
            post_head<A>(k, hTA, slope, 0, pTA);
            if (act) post_head<A>(k, hA, slope, A * RB, pA);
            {
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r, b = b0 + m;
                    if (m >= RB) continue;
                    float ap = 0.f, pi = 0.f;   // lane j < 4 keeps output j
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float nz = fminf(fmaxf(st.noise[j], -c.noise_clamp), c.noise_clamp);                       // HIRL.py:266
                        const float a4 = fminf(fmaxf(tanhf(get_head<A>(xb, hTA, 0, m, j)) + nz, -1.f), 1.f);                // HIRL.py:267
                        const float p4 = act ? tanhf(get_head<A>(xb, hA, A * RB, m, j)) : 0.f;
                        if (lane == j) { ap = a4; pi = p4; }
                    }
                    if (lane < A) {
                        k.sm[O_AP + m * A + lane] = ap;
                        if (k.crank == 0 && b < B) {
                            c.sc.APRIME[(size_t)b * A + lane] = ap;
                            if (act) c.sc.PI[(size_t)b * A + lane] = pi;
                        }
                    }
                    if (act && b < B) {   // forward state of pass A for the actor backward in phase C
                        c.sc.XH2[(size_t)b * H2 + NS * k.crank + lane] = pA.z[r][0];
                        c.sc.XH2[(size_t)b * H2 + NS * k.crank + lane + 32] = pA.z[r][1];
                        if (k.crank == 0 && lane == 0) c.sc.RS2[b] = pA.rstd[r];
                    }
                }
            }
            if (act && hirl) {
                x_begin(k, 2 * A * RB * CL);
                post_head<A>(k, hA, slope, 0, pAB);
                post_head<A>(k, hBC, slope, A * RB, pBC);
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r, b = b0 + m;
                    if (m >= RB) continue;
                    float pib = 0.f, abc = 0.f;
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float v0 = tanhf(get_head<A>(xb, hA, 0, m, j)), v1 = tanhf(get_head<A>(xb, hBC, A * RB, m, j));
                        if (lane == j) { pib = v0; abc = v1; }
                    }
                    if (b < B) {
                        if (k.crank == 0 && lane < A) { c.sc.PIB[(size_t)b * A + lane] = pib; c.sc.ABC[(size_t)b * A + lane] = abc; }
                        const size_t row = (size_t)c.Bpad + b;
                        c.sc.XH2[row * H2 + NS * k.crank + lane] = pAB.z[r][0];
                        c.sc.XH2[row * H2 + NS * k.crank + lane + 32] = pAB.z[r][1];
                        if (k.crank == 0 && lane == 0) c.sc.RS2[row] = pAB.rstd[r];
                    }
                }
            }
            __syncthreads();   // a' of every row is in shared memory
            // ---- first layers of the target critics on (s', a')  (HIRL.py:270)
            for (int inst = warp; inst < 2 * RPW; inst += WARPS) {
                const int t = inst / RPW, q = inst - t * RPW, m = k.crank + 8 * q;
                if (m >= RB) continue;
                const int lanei = threadIdx.x & 31;
                (void)lanei;
                float x[SA];
#pragma unroll
                for (int kk = 0; kk < S; ++kk) x[kk] = k.sm[O_XROW + m * 32 + SA + kk];
#pragma unroll
                for (int j = 0; j < A; ++j) x[S + j] = k.sm[O_AP + m * A + j];
                l1_row<CriticL>(c.tcritic + t * CriticL::SIZE, x, slope, lane, c.sc.H1P + ((size_t)(P_TC1 + t) * c.Bpad + b0 + m) * WLD, nullptr, nullptr);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            // ---- segment 2: TC1, TC2, C1, C2
            Pass pT[2], pC[2];
            lda(k, c.sc.H1P + P_TC1 * pstride + rowoff);
            lda(k, c.sc.H1P + P_TC2 * pstride + rowoff);
            gemm(k, true);
            read_z(k, hT[0], pT[0]);
            ldw(k, c.w2p_critic);
            lda(k, c.sc.H1P + P_C1 * pstride + rowoff);
            gemm(k, true);
            read_z(k, hT[1], pT[1]);
            ldw(k, c.w2p_critic + (size_t)H2 * WLD);
            lda(k, c.sc.H1P + P_C2 * pstride + rowoff);
            const float *wC1 = k.sm + (k.wo & 1u) * W_FLOATS;
            gemm(k, true);
            read_z(k, hC[0], pC[0]);
            const float *wC2 = k.sm + (k.wo & 1u) * W_FLOATS;
            gemm(k, true);
            read_z(k, hC[1], pC[1]);
            x_begin(k, 4 * 2 * RB * CL);
            post_stats(k, 0, pT[0]); post_stats(k, 1, pT[1]); post_stats(k, 2, pC[0]); post_stats(k, 3, pC[1]);
            {
                const float *xb = x_wait(k);
                finish_stats(k, xb, 0, pT[0]); finish_stats(k, xb, 1, pT[1]); finish_stats(k, xb, 2, pC[0]); finish_stats(k, xb, 3, pC[1]);
            }
            x_begin(k, 4 * RB * CL);
            post_head<1>(k, hT[0], slope, 0, pT[0]); post_head<1>(k, hT[1], slope, RB, pT[1]);
            post_head<1>(k, hC[0], slope, 2 * RB, pC[0]); post_head<1>(k, hC[1], slope, 3 * RB, pC[1]);
            float dq[2][RPW][4];
            {
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r, b = b0 + m;
                    dq[0][r][0] = dq[1][r][0] = 0.f;
                    if (m >= RB) continue;
                    const float t1 = get_head<1>(xb, hT[0], 0, m, 0), t2 = get_head<1>(xb, hT[1], RB, m, 0);
                    const float q1 = get_head<1>(xb, hC[0], 2 * RB, m, 0), q2 = get_head<1>(xb, hC[1], 3 * RB, m, 0);
                    if (b >= B) continue;
                    const float rew = k.sm[O_XROW + m * 32 + 30], d = k.sm[O_XROW + m * 32 + 31];
                    const float tq = fminf(t1, t2);
                    const float y = __fadd_rn(rew, __fmul_rn(__fmul_rn(c.gamma, tq), __fsub_rn(1.f, d)));   // HIRL.py:274
                    const float d1 = q1 - y, d2 = q2 - y;
                    loss[0] = fmaf(d1, d1, loss[0]);
                    loss[1] = fmaf(d2, d2, loss[1]);
                    dq[0][r][0] = (2.f / (float)B) * d1;   // mse_loss backward, mean reduction (HIRL.py:284)
                    dq[1][r][0] = (2.f / (float)B) * d2;
                    if (k.crank == 0 && lane == 0) {
                        c.sc.Q[b] = q1; c.sc.Q[(size_t)B + b] = q2; c.sc.TQ[b] = t1; c.sc.TQ[(size_t)B + b] = t2; c.sc.Y[b] = y;
                    }
                }
            }
            // ---- backward through head + LayerNorm2 of both towers
            Bwd bw[2];
            x_begin(k, 2 * 2 * RB * CL);
            ln2_bwd_post<1>(k, hC[0], slope, 0, pC[0], dq[0], bw[0]);
            ln2_bwd_post<1>(k, hC[1], slope, 2 * RB, pC[1], dq[1], bw[1]);
            {
                const float *xb = x_wait(k);
                ln2_bwd_finish(xb, 0, pC[0], bw[0]);
                ln2_bwd_finish(xb, 2 * RB, pC[1], bw[1]);
            }
            // column sums over the block's rows: db2, dg2, dbe2, dW3 of each tower (8 quantities x 64 columns)
            __syncthreads();
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int n0 = NS * k.crank + lane;
                const float g0 = hC[t].g2[n0], g1 = hC[t].g2[n0 + 32], e0 = hC[t].be2[n0], e1 = hC[t].be2[n0 + 32];
                float v[RPW][2];
                stage_col(k, 4 * t + 0, 8, bw[t].dxh);
#pragma unroll
                for (int r = 0; r < RPW; ++r) { v[r][0] = bw[t].dy[r][0] * pC[t].z[r][0]; v[r][1] = bw[t].dy[r][1] * pC[t].z[r][1]; }
                stage_col(k, 4 * t + 1, 8, v);
                stage_col(k, 4 * t + 2, 8, bw[t].dy);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    v[r][0] = dq[t][r][0] * act_f(fmaf(pC[t].z[r][0], g0, e0), slope);
                    v[r][1] = dq[t][r][0] * act_f(fmaf(pC[t].z[r][1], g1, e1), slope);
                }
                stage_col(k, 4 * t + 3, 8, v);
                if (lane == 0) {
                    float s = 0.f;
#pragma unroll
                    for (int r = 0; r < RPW; ++r) s += dq[t][r][0];
                    k.sm[O_MISC + warp * 2 + t] = s;
                }
            }
            __syncthreads();
            col[0] += take_col(k, threadIdx.x, 8);
            col[1] += take_col(k, threadIdx.x + CTHREADS, 8);
            if (threadIdx.x < 2) {
                float s = 0.f;
                for (int w = 0; w < WARPS; ++w) s += k.sm[O_MISC + w * 2 + threadIdx.x];
                db3[threadIdx.x & 1] += s;
            }
            __syncthreads();
            store_dz(k, c, 0, 0, b0, bw[0]);
            store_dz(k, c, 1, 1, b0, bw[1]);
            dh1_two(k, wC1, wC2, true);
            for (int inst = warp; inst < 2 * RPW; inst += WARPS) {
                const int t = inst / RPW, q = inst - t * RPW;
                ln1_rows<CriticL>(k, c, c.critic + t * CriticL::SIZE, t, t, b0, slope, q, nullptr);
            }
        }
        // ---- per-cluster partial sums -> global
        {
            float *part = c.sc.PART2C + (size_t)k.cid * PC_WORDS;
            const int q = threadIdx.x >> 6, n = NS * k.crank + (threadIdx.x & 63);
            part[(0 * 4 + q) * H2 + n] = col[0];
            part[(1 * 4 + q) * H2 + n] = col[1];
            if (k.crank == 0 && threadIdx.x < 2) part[2 * 4 * H2 + threadIdx.x] = db3[threadIdx.x & 1];
            flush_scalar(k, c, 0, loss[0]);
            flush_scalar(k, c, 1, loss[1]);
        }
    }

    // ===========================================================================================================
    // phase C (modes 0 / 1, actor calls): Q1 with the UPDATED critic, soft count, d pi, actor backward
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_c(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        const bool hirl = st.mode == 0, soft = hirl && st.bc_weight_now == 100.f;
        const Head hQ = head_of<CriticL>(c.critic);
        ColAcc ca;
        ca.a[0][0] = ca.a[0][1] = ca.a[1][0] = ca.a[1][1] = 0.f;
        ca.db3[0] = ca.db3[1] = 0.f; ca.bc_sq = ca.fire_sq = 0.f;
        float cnt = 0.f, sum_q = 0.f;
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_critic);
            gather_rows(k, c, st, step, b0, true, hirl);
            __syncthreads();
            for (int inst = warp; inst < (soft ? 2 : 1) * RPW; inst += WARPS) {
                const int pi = inst / RPW, q = inst - pi * RPW, m = k.crank + 8 * q, b = b0 + m;
                if (m >= RB) continue;
                float x[SA];
#pragma unroll
                for (int kk = 0; kk < S; ++kk) x[kk] = k.sm[O_XROW + m * 32 + kk];
                const float *act = (pi == 0 ? c.sc.PI : c.sc.ABC) + (size_t)min(b, B - 1) * A;
#pragma unroll
                for (int j = 0; j < A; ++j) x[S + j] = b < B ? act[j] : 0.f;
                const bool keep = pi == 0 && b < B;
                l1_row<CriticL>(c.critic, x, slope, lane, c.sc.H1P + ((size_t)(pi == 0 ? P_QP : P_QS) * c.Bpad + b) * WLD,
                                keep ? XH1p(c, 4) + (size_t)b * H1 : nullptr, keep ? RS1p(c, 4) + b : nullptr);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            const size_t rowoff = (size_t)b0 * WLD, pstride = (size_t)c.Bpad * WLD;
            lda(k, c.sc.H1P + P_QP * pstride + rowoff);
            if (soft) lda(k, c.sc.H1P + P_QS * pstride + rowoff);
            ldw(k, c.w2p_actor);
            Pass pQ, pS;
            const float *wQ = k.sm + (k.wo & 1u) * W_FLOATS;
            gemm(k, !soft);
            read_z(k, hQ, pQ);
            if (soft) { gemm(k, true); read_z(k, hQ, pS); }
            const float *wAct = k.sm + (k.wo & 1u) * W_FLOATS;
            x_begin(k, (soft ? 2 : 1) * 2 * RB * CL);
            post_stats(k, 0, pQ);
            if (soft) post_stats(k, 1, pS);
            {
                const float *xb = x_wait(k);
                finish_stats(k, xb, 0, pQ);
                if (soft) finish_stats(k, xb, 1, pS);
            }
            x_begin(k, (soft ? 2 : 1) * RB * CL);
            post_head<1>(k, hQ, slope, 0, pQ);
            if (soft) post_head<1>(k, hQ, slope, RB, pS);
            float dq[RPW][4];
            {
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r, b = b0 + m;
                    dq[r][0] = 0.f;
                    if (m >= RB) continue;
                    const float q = get_head<1>(xb, hQ, 0, m, 0);
                    const float sq = soft ? get_head<1>(xb, hQ, RB, m, 0) : 0.f;
                    if (b >= B) continue;
                    dq[r][0] = -1.f / (float)B;          // rl_loss = -mean(Q1)  (HIRL.py:297)
                    sum_q += q;
                    if (soft) cnt += sq > q ? 1.f : 0.f;   // HIRL.py:303
                    if (k.crank == 0 && lane == 0) { c.sc.RLQ[b] = q; if (soft) c.sc.SOFTQ[b] = sq; }
                }
            }
            Bwd bw;
            x_begin(k, 2 * RB * CL);
            ln2_bwd_post<1>(k, hQ, slope, 0, pQ, dq, bw);
            {
                const float *xb = x_wait(k);
                ln2_bwd_finish(xb, 0, pQ, bw);
            }
            __syncthreads();
            {   // shared copy only: the critic's own dW2 from the actor loss is discarded (HIRL.py:286 zero_grad)
                float *dzs = k.sm + O_DZS;
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r;
                    if (m >= RB) continue;
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) dzs[((lane + 32 * cc) * 2 + (m & 1)) * RIP + (m >> 1)] = bw.dxh[r][cc];
                }
            }
            dh1_two(k, wQ, wQ, false);
            // LayerNorm1 backward of Q1(s, pi) on the rows this CTA owns -> d pi, shared with the whole cluster
            x_begin(k, A * RB);
            for (int q = warp; q < RPW; q += WARPS) {
                float g4[A];
                ln1_rows<CriticL>(k, c, c.critic, 0, 4, b0, slope, q, g4, false);
                const int m = k.crank + 8 * q;
                if (m < RB) {
#pragma unroll
                    for (int j = 0; j < A; ++j) x_post(k, m * A + j, g4[j]);
                    if (b0 + m < B && lane < A) {
                        float v = g4[0];
#pragma unroll
                        for (int j = 1; j < A; ++j) if (lane == j) v = g4[j];
                        c.sc.GPI[(size_t)(b0 + m) * A + lane] = v;
                    }
                }
            }
            float gpi[RPW][4];
            {
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r;
#pragma unroll
                    for (int j = 0; j < A; ++j) gpi[r][j] = m < RB ? xb[(m & 7) * XS + m * A + j] : 0.f;
                }
            }
            actor_backward(k, c, st, b0, true, hirl, gpi, wAct, ca);
        }
        flush_actor_cols(k, c, ca, true, hirl);
        flush_scalar(k, c, 2, cnt);
        flush_scalar(k, c, 3, sum_q);
        flush_scalar(k, c, 4, ca.bc_sq);
        flush_scalar(k, c, 5, ca.fire_sq);
    }

    // ===========================================================================================================
    // behaviour cloning (mode 2, BC.py:160-185): expert-state actor pass + its backward
    // ===========================================================================================================
    static __device__ __forceinline__ void phase_bc(Ctl &k, const AgentCtx &c, const hirl_step_t &st, int step) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B, nblk = (B + RB - 1) / RB;
        const float slope = slope_of(c);
        const Head hA = head_of<ActorL>(c.actor);
        ColAcc ca;
        ca.a[0][0] = ca.a[0][1] = ca.a[1][0] = ca.a[1][1] = 0.f;
        ca.db3[0] = ca.db3[1] = 0.f; ca.bc_sq = ca.fire_sq = 0.f;
        for (int blk = k.cid; blk < nblk; blk += k.ncl) {
            const int b0 = blk * RB;
            ldw(k, c.w2p_actor);
            gather_rows(k, c, st, step, b0, false, true);
            __syncthreads();
            for (int q = warp; q < RPW; q += WARPS) {
                const int m = k.crank + 8 * q;
                if (m < RB) l1_inst<ActorL>(c, c.actor, k.sm + O_XBROW + m * BCW, P_AB, b0 + m, slope);
            }
            fence_proxy_async_all();
            cluster_sync_all();
            lda(k, c.sc.H1P + ((size_t)P_AB * c.Bpad + b0) * WLD);
            Pass p;
            const float *wAct = k.sm + (k.wo & 1u) * W_FLOATS;
            mbar_wait(k.wbar + (k.wo & 1u), (k.wo >> 1) & 1u, 5);
            mbar_wait(k.abar + (k.ao & 1u), (k.ao >> 1) & 1u, 6);
            fwd_pass<RI>(k.sm + O_A0 + (k.ao & 1u) * RB * WLD, wAct, k.sm + O_RED);
            k.ao += 1u;   // the weight slice stays current for the backward pass (consumed in actor_backward)
            read_z(k, hA, p);
            x_begin(k, 2 * RB * CL);
            post_stats(k, 0, p);
            {
                const float *xb = x_wait(k);
                finish_stats(k, xb, 0, p);
            }
            x_begin(k, A * RB * CL);
            post_head<A>(k, hA, slope, 0, p);
            {
                const float *xb = x_wait(k);
#pragma unroll
                for (int r = 0; r < RPW; ++r) {
                    const int m = warp + 8 * r, b = b0 + m;
                    if (m >= RB) continue;
                    float pib = 0.f;
#pragma unroll
                    for (int j = 0; j < A; ++j) {
                        const float v0 = tanhf(get_head<A>(xb, hA, 0, m, j));
                        if (lane == j) pib = v0;
                    }
                    if (b < B) {
                        if (lane < A) c.sc.PIB[(size_t)b * A + lane] = pib;   // every CTA writes the same value: read back below by this CTA
                        const size_t row = (size_t)c.Bpad + b;
                        c.sc.XH2[row * H2 + NS * k.crank + lane] = p.z[r][0];
                        c.sc.XH2[row * H2 + NS * k.crank + lane + 32] = p.z[r][1];
                        c.sc.RS2[row] = p.rstd[r];
                    }
                }
            }
            __syncthreads();
            float gpi[RPW][4];
#pragma unroll
            for (int r = 0; r < RPW; ++r)
#pragma unroll
                for (int j = 0; j < A; ++j) gpi[r][j] = 0.f;
            actor_backward(k, c, st, b0, false, true, gpi, wAct, ca);
        }
        flush_actor_cols(k, c, ca, false, true);
        flush_scalar(k, c, 4, ca.bc_sq);
        flush_scalar(k, c, 5, ca.fire_sq);
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
    static __device__ __forceinline__ int nclu_of(const Ctl &k, const AgentCtx &c) {
        const int nblk = (c.B + RB - 1) / RB;
        return nblk < k.ncl ? nblk : k.ncl;
    }
    // first-layer parameter gradients of one tower, one output column per warp: db1, dg1, dbe1, dW1[k]  (two optional
    // branches combined with weights wa, wb) followed by Adam on those k0 + 3 parameters
    template <class L>
    static __device__ __forceinline__ void l1_columns(const Ctl &k, const AgentCtx &c, int ntowers, int slot_a, int slot_b, float wa, float wb, const float *xa, int xlda,
                                                      const float *xb_, int xldb, float *w, float *gw, float *m, float *v, float *tgt, const Opt &o) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, B = c.B;
        for (int job = blockIdx.x * WARPS + warp; job < ntowers * H1; job += gridDim.x * WARPS) {
            const int t = job / H1, n = job - t * H1;
            float tot[L::k0 + 3];
#pragma unroll
            for (int i = 0; i < L::k0 + 3; ++i) tot[i] = 0.f;
#pragma unroll
            for (int br = 0; br < 2; ++br) {
                const int slot = br ? slot_b : slot_a + t;
                if (slot < 0) continue;
                const float wt = br ? wb : wa;
                const float *xs = br ? xb_ : xa;
                const int xld = br ? xldb : xlda;
                float acc[L::k0 + 3];
#pragma unroll
                for (int i = 0; i < L::k0 + 3; ++i) acc[i] = 0.f;
                const float *dz1 = c.sc.DZ1 + (size_t)slot * B * H1 + n, *dy1 = c.sc.DY1 + (size_t)slot * B * H1 + n, *xh1 = XH1p(c, slot) + n;
                for (int b = lane; b < B; b += 32) {
                    const float dz = __ldcg(dz1 + (size_t)b * H1), dy = __ldcg(dy1 + (size_t)b * H1), xh = __ldcg(xh1 + (size_t)b * H1);
                    acc[L::k0] += dz;
                    acc[L::k0 + 1] = fmaf(dy, xh, acc[L::k0 + 1]);
                    acc[L::k0 + 2] += dy;
#pragma unroll
                    for (int kk = 0; kk < L::k0; ++kk) acc[kk] = fmaf(dz, __ldcg(xs + (size_t)b * xld + kk), acc[kk]);
                }
#pragma unroll
                for (int i = 0; i < L::k0 + 3; ++i) tot[i] = fmaf(wt, warp_sum(acc[i]), tot[i]);
            }
            float mine = 0.f;
#pragma unroll
            for (int i = 0; i < L::k0 + 3; ++i)
                if (lane == i) mine = tot[i];
            if (lane < L::k0 + 3) {
                const int off = lane < L::k0 ? L::W1T + lane * H1 + n : lane == L::k0 ? L::B1 + n : lane == L::k0 + 1 ? L::G1 + n : L::BE1 + n;
                apply(mine, w, gw, m, v, tgt, o, t * L::SIZE + off, nullptr, nullptr, 0);
            }
        }
    }
    // dW2 tiles of `ntowers` towers; per tower up to two row sets (branches) combined with weights wa, wb
    template <class L, int TK>
    static __device__ __forceinline__ void dw2_tiles(const Ctl &k, const AgentCtx &c, int ntowers, int slot_a, int slot_b, int pass_a, int pass_b, float wa, float wb,
                                                     float *w, float *gw, float *m, float *v, float *tgt, float *img, float *timg, const Opt &o) {
        constexpr int KPT = TK / 4, NKT = H1 / TK;
        float *Dz = k.sm, *Hs = k.sm + 256 * NS, *red = Hs + 256 * KS;
        const int B = c.B, tid = threadIdx.x;
        for (int tile = blockIdx.x; tile < ntowers * CL * NKT; tile += gridDim.x) {
            const int t = tile / (CL * NKT), rem = tile - t * CL * NKT, ns = rem / NKT, k0 = (rem - ns * NKT) * TK;
            float g[KPT];
#pragma unroll
            for (int j = 0; j < KPT; ++j) g[j] = 0.f;
#pragma unroll
            for (int br = 0; br < 2; ++br) {
                const int slot = br ? slot_b : slot_a + t, pass = br ? pass_b : pass_a + t;
                if (slot < 0) continue;
                const float wt = br ? wb : wa;
                float acc[8][KPT];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < KPT; ++j) acc[i][j] = 0.f;
                const float *dzs = c.sc.DZ2S + ((size_t)(slot * CL + ns) * c.Bpad) * NS;
                const float *h1p = c.sc.H1P + (size_t)pass * c.Bpad * WLD;
                for (int c0 = 0; c0 < B; c0 += 256) {
                    const int rows = min(256, B - c0), rp = (rows + 7) & ~7;
                    __syncthreads();
                    dw_stage<TK>(Dz, Hs, dzs, h1p, k0, c0, B, rp);
                    cp_commit();
                    cp_wait_all();
                    __syncthreads();
                    dw_accumulate<TK>(Dz, Hs, rp, acc);
                }
                float gb[KPT];
                dw_reduce<TK>(acc, red, gb);
#pragma unroll
                for (int j = 0; j < KPT; ++j) g[j] = fmaf(wt, gb[j], g[j]);
            }
            const int n = NS * ns + (tid & 63), kk = k0 + KPT * (tid >> 6);
#pragma unroll
            for (int j = 0; j < KPT; ++j)
                apply(g[j], w, gw, m, v, tgt, o, t * L::SIZE + L::W2 + n * H1 + kk + j, img, timg, (t * H2 + n) * WLD + kk + j);
        }
    }

    static __device__ __forceinline__ void phase_b(Ctl &k, const AgentCtx &c, const hirl_step_t &st, bool tk16) {
        const int nclu = nclu_of(k, c);
        Opt o{st.critic_neg_step, st.critic_bc2_sqrt, false, 0.f, 0.f};
        // small second-layer parameters: fixed-order sum of the per-cluster partials, then Adam
        for (int e = blockIdx.x * CTHREADS + threadIdx.x; e < 2 * 4 * H2 + 2; e += gridDim.x * CTHREADS) {
            float g = 0.f;
            for (int cc = 0; cc < nclu; ++cc) g += __ldcg(c.sc.PART2C + (size_t)cc * PC_WORDS + e);
            int off;
            if (e < 2 * 4 * H2) {
                const int t = e / (4 * H2), q = (e / H2) & 3, n = e & (H2 - 1);
                off = t * CriticL::SIZE + (q == 0 ? CriticL::B2 : q == 1 ? CriticL::G2 : q == 2 ? CriticL::BE2 : CriticL::W3) + n;
            } else {
                off = (e - 2 * 4 * H2) * CriticL::SIZE + CriticL::B3;
            }
            apply(g, c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, o, off, nullptr, nullptr, 0);
        }
        l1_columns<CriticL>(k, c, 2, 0, -1, 1.f, 0.f, c.sc.X, ROW, nullptr, 0, c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, o);
        if (blockIdx.x == 0 && threadIdx.x < 32) {   // critic_loss = mse(Q1,y) + mse(Q2,y)  (HIRL.py:284)
            float s0 = 0.f, s1 = 0.f;
            for (int cc = 0; cc < nclu; ++cc) { s0 += __ldcg(c.sc.LP + cc * LPW); s1 += __ldcg(c.sc.LP + cc * LPW + 1); }
            if (threadIdx.x == 0) c.carry[0] = __fadd_rn(s0 / (float)c.B, s1 / (float)c.B);
        }
        if (tk16) dw2_tiles<CriticL, 16>(k, c, 2, 0, -1, P_C1, 0, 1.f, 0.f, c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, c.w2p_critic, nullptr, o);
        else dw2_tiles<CriticL, 32>(k, c, 2, 0, -1, P_C1, 0, 1.f, 0.f, c.critic, c.g_critic, c.m_critic, c.v_critic, nullptr, c.w2p_critic, nullptr, o);
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
        Opt o{st.actor_neg_step, st.actor_bc2_sqrt, st.do_polyak != 0, c.tau, (float)(1.0 - (double)c.tau)};
        for (int e = blockIdx.x * CTHREADS + threadIdx.x; e < 7 * H2 + A; e += gridDim.x * CTHREADS) {
            float ga = 0.f, gb = 0.f;
            const int ea = e < 7 * H2 ? e : 2 * 7 * H2 + (e - 7 * H2), eb = e < 7 * H2 ? 7 * H2 + e : 2 * 7 * H2 + 4 + (e - 7 * H2);
            for (int cc = 0; cc < nclu; ++cc) {
                if (rl) ga += __ldcg(c.sc.PART2A + (size_t)cc * PA_WORDS + ea);
                if (bc) gb += __ldcg(c.sc.PART2A + (size_t)cc * PA_WORDS + eb);
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
        l1_columns<ActorL>(k, c, 1, rl ? 2 : -1, bc ? 3 : -1, wrl, wbc, c.sc.X, ROW, c.sc.XB, BCW, c.actor, c.g_actor, c.m_actor, c.v_actor, c.tactor, o);
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
        if (blockIdx.x == 0 && threadIdx.x == 0) {
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
        const int sa = rl ? 2 : -1, sb2 = bc ? 3 : -1;
        if (tk16) dw2_tiles<ActorL, 16>(k, c, 1, sa, sb2, P_A, P_AB, wrl, wbc, c.actor, c.g_actor, c.m_actor, c.v_actor, c.tactor, c.w2p_actor, c.w2p_tactor, o);
        else dw2_tiles<ActorL, 32>(k, c, 1, sa, sb2, P_A, P_AB, wrl, wbc, c.actor, c.g_actor, c.m_actor, c.v_actor, c.tactor, c.w2p_actor, c.w2p_tactor, o);
    }
};

// ---------------------------------------------------------------------------------------------------------------
template <int RI>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CTHREADS, 1) hirl_cluster_kernel(const __grid_constant__ ClusterLaunch L) {
    using K = Ker<RI>;
    extern __shared__ float4 smem4[];
    __shared__ AgentCtx ctx;
    __shared__ hirl_step_t st;
    __shared__ uint64_t bars[6];
    {
        const int *src = reinterpret_cast<const int *>(&L.agent);
        int *dst = reinterpret_cast<int *>(&ctx);
        for (int i = threadIdx.x; i < (int)(sizeof(AgentCtx) / sizeof(int)); i += CTHREADS) dst[i] = src[i];
    }
    typename K::Ctl k;
    k.sm = reinterpret_cast<float *>(smem4);
    k.wbar = bars; k.abar = bars + 2; k.xbar = bars + 4;
    k.crank = (int)cluster_ctarank(); k.cid = (int)cluster_id_x(); k.ncl = L.ncl;
    k.wi = k.wo = k.ai = k.ao = 0u; k.xr = 0u;
    k.flags = L.bar_flags;
    k.gen = ld_relaxed_u32(k.flags + k.cid * FLAG_STRIDE);
    if (threadIdx.x == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 2 * NS * 2 * K::RIP; i += CTHREADS) k.sm[K::O_DZS + i] = 0.f;   // padding rows of the dz2 copies stay zero
    __syncthreads();
    cluster_sync_all();   // peers' barriers are initialised before anybody posts to them
    const bool tk16 = (int)gridDim.x < 2 * CL * (H1 / 32);
    long long *prof = L.prof;
    for (int step = 0; step < L.k_steps; ++step) {
        __syncthreads();
        if (threadIdx.x < (int)(sizeof(hirl_step_t) / sizeof(int)))
            reinterpret_cast<int *>(&st)[threadIdx.x] = reinterpret_cast<const int *>(ctx.steps + step)[threadIdx.x];
        __syncthreads();
        long long *pr = (prof && step < PROF_STEPS && blockIdx.x == 0 && threadIdx.x == 0) ? prof + (size_t)step * 16 : nullptr;
        if (pr) pr[0] = clock64();
        const bool bc_only = st.mode == 2;
#define HIRL_PH(i, call)                  \
    call;                                 \
    if (pr) pr[2 * (i) + 1] = clock64();  \
    K::grid_sync(k);                      \
    if (pr) pr[2 * (i) + 2] = clock64();
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
        if (blockIdx.x == 0 && threadIdx.x < HIRL_OUT_SCALARS) ctx.out_scalars[(size_t)step * HIRL_OUT_SCALARS + threadIdx.x] = ctx.carry[threadIdx.x];
    }
}

}  // namespace cl
}  // namespace hirl
