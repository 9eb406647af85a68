// cluster_gemm.cuh -- building blocks of the cluster-resident update kernel (cluster_update.cuh).
//
// A thread-block cluster of CL = 8 CTAs owns a block of RB = 16 batch rows through every layer of every network.
// CTA r of the cluster holds the weight slice W2[64 r .. 64 r + 63][0..255] of the network in use in shared memory
// (streamed from L2 by cp.async.bulk, double buffered, 65 KB per slice) and computes the 64 output features of that
// slice for all 16 rows; row statistics (LayerNorm2, heads) are completed across the cluster through distributed
// shared memory.  FP32 FMA throughout (parity with torch fp32 matmuls, reference hirl/agents/HIRL.py:55-97,128-140).
#pragma once
#include "common.cuh"
#include "tc_gemm.cuh"   // mbarrier helpers (tc::mbar_*)

namespace hirl {
namespace cl {

using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_u32;

constexpr int CL = 8;            // CTAs per cluster
constexpr int NS = H2 / CL;      // 64 output features (W2 rows) per CTA
constexpr int KS = H1 / CL;      // 32 hidden-1 columns per CTA (dW2 tiles)
constexpr int CTHREADS = 256;
constexpr int WLD = H1 + 4;      // padded row stride (floats) of weight slices and activation blocks: rows 1040 B apart
constexpr int W_FLOATS = NS * WLD;   // 16 640 floats = 66 560 B per weight slice

// ---- cluster / DSMEM primitives --------------------------------------------------------------------------------
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
// all threads of all CTAs of the cluster; release/acquire at cluster scope (orders global and shared::cluster accesses)
__device__ __forceinline__ void cluster_sync_all() {
    __syncwarp();   // the .aligned barrier needs a converged warp (see team_sync in update.cuh)
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// remote store that signals the DESTINATION CTA's mbarrier (4 transaction bytes): DSMEM exchange without a cluster barrier
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(remote_addr), "r"(__float_as_uint(v)),
                 "r"(remote_bar)
                 : "memory");
}

// ---- bulk asynchronous copies (TMA, non-tensor form) global -> shared, completion on an mbarrier ---------------
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// generic-proxy writes (st.global by any CTA, made visible to this thread by a barrier) -> later async-proxy reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- forward pass of one row block: z[m][n] = sum_k A[m][k] W[n][k]  (RB x 64 x 256, RB = 2 RI) ------------------
// As [RB][WLD], Ws [64][WLD] in shared memory.  K is split over the 8 warps (32 k each); a lane owns rows mg + 2 i
// (mg = lane>>4, i < RI) and slice columns ng + 16 j (ng = lane&15, j < 4).  Measured on B200 (tools/cluster_probe3.cu):
// an LDS.128 costs 4 cycles of the shared-memory pipe when a quarter warp reads 8 distinct addresses (weights) and
// 2.5 when it reads one (activations, broadcast) whether or not the other quarters repeat them, so tall lane tiles
// (many activation rows, few weight columns) are the cheap ones: RI + 4 LDS.128 per 16 RI FFMA.
// The 8 partial tiles are summed through red[4][RB][RL] in a fixed order (deterministic): after the call the value of
// (m, n) is (red[0]+red[1]) + (red[2]+red[3]) at [m * RLD + n].  All 256 threads; contains __syncthreads.
template <int RI>
struct RedLd { static constexpr int v = RI == 8 ? 80 : 64; };   // row stride of the reduction buffers (80: lanes mg = 0 / 1 land 16 banks apart)
template <int RI>
__device__ __forceinline__ void fwd_round(const float *ap, const float *bp, int r, float4 (&a)[RI], float4 (&b)[4]) {
#pragma unroll
    for (int i = 0; i < RI; ++i) a[i] = *reinterpret_cast<const float4 *>(ap + 2 * i * WLD + 4 * r);
#pragma unroll
    for (int j = 0; j < 4; ++j) b[j] = *reinterpret_cast<const float4 *>(bp + 16 * j * WLD + 4 * r);
}
template <int RI>
__device__ __forceinline__ void fwd_fma(const float4 (&a)[RI], const float4 (&b)[4], float (&acc)[RI][4]) {
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
            acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
            acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
            acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
        }
}
template <int RI>
__device__ __forceinline__ void fwd_pass(const float *As, const float *Ws, float *red) {
    constexpr int RBv = 2 * RI, RL = RedLd<RI>::v;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mg = lane >> 4, ng = lane & 15;
    const float *ap = As + mg * WLD + 32 * warp;
    const float *bp = Ws + ng * WLD + 32 * warp;
    float acc[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float4 a0[RI], b0[4], a1[RI], b1[4];
    fwd_round<RI>(ap, bp, 0, a0, b0);
    // rolled (instruction-cache resident): 4 iterations of two k-rounds with ping-pong fragments
#pragma unroll 1
    for (int r = 0; r < 8; r += 2) {
        fwd_round<RI>(ap, bp, r + 1, a1, b1);
        fwd_fma<RI>(a0, b0, acc);
        if (r + 2 < 8) fwd_round<RI>(ap, bp, r + 2, a0, b0);
        fwd_fma<RI>(a1, b1, acc);
    }
    float *p = red + (warp & 3) * RBv * RL + mg * RL + ng;
    __syncthreads();   // readers of the previous pass' sums are done
    if (warp >= 4) {
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) p[2 * i * RL + 16 * j] = acc[i][j];
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int i = 0; i < RI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) p[2 * i * RL + 16 * j] += acc[i][j];
    }
    __syncthreads();
}
// ---- the same pass on the tensor cores: mma.sync m16n8k8 TF32 with the 3-product split (3xTF32) ------------------
// x = hi + lo, hi = x with the 13 low mantissa bits cleared (what the tensor core reads from an fp32 word), lo = x - hi
// (exact in fp32); z += lo.hi + hi.lo + hi.hi, accumulated in fp32.  The K range of a warp is only 32, so an accumulation
// chain is 12 MMAs long and the 8 warps' partial tiles are added in fp32 as before: the result is as close to the fp64
// product as the FFMA version (tools/cluster_probe4.cu).  Fragments are read straight from the padded [rows][260]
// operand layouts with conflict-free LDS.32 (bank = 4 row + k), a quarter of the shared-memory traffic of the FFMA tile.
// Rows 16 .. RB-1 (RB = 18) do not fill a second m16 tile; they are computed with FFMA by all threads.
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int RI>
__device__ __forceinline__ void fwd_pass_mma(const float *As, const float *Ws, float *red) {
    constexpr int RBv = 2 * RI, RL = RedLd<RI>::v;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const float *ap = As + g * WLD + 32 * warp + t, *bp = Ws + g * WLD + 32 * warp + t;
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {
        uint32_t ahi[4], alo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float x = ap[(i & 1) * 8 * WLD + (i >> 1) * 4 + 8 * ks];
            ahi[i] = __float_as_uint(x) & 0xffffe000u;
            alo[i] = __float_as_uint(x - __uint_as_float(ahi[i]));
        }
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            const float y0 = bp[nt * 8 * WLD + 8 * ks], y1 = bp[nt * 8 * WLD + 8 * ks + 4];
            const uint32_t h0 = __float_as_uint(y0) & 0xffffe000u, h1 = __float_as_uint(y1) & 0xffffe000u;
            const uint32_t l0 = __float_as_uint(y0 - __uint_as_float(h0)), l1 = __float_as_uint(y1 - __uint_as_float(h1));
            mma_tf32(acc[nt], alo, h0, h1);
            mma_tf32(acc[nt], ahi, l0, l1);
            mma_tf32(acc[nt], ahi, h0, h1);
        }
    }
    // rows 16.. (RB = 18): FFMA, thread = (row, column quarter) x k-range of the warp
    float ex[4] = {0.f, 0.f, 0.f, 0.f};
    if (RBv > 16) {
        const int er = lane >> 4, ec = lane & 15;
        const float *eap = As + (16 + er) * WLD + 32 * warp, *ebp = Ws + ec * WLD + 32 * warp;
#pragma unroll 2
        for (int r = 0; r < 8; ++r) {
            const float4 a = *reinterpret_cast<const float4 *>(eap + 4 * r);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 b = *reinterpret_cast<const float4 *>(ebp + 16 * j * WLD + 4 * r);
                ex[j] = fmaf(a.x, b.x, ex[j]); ex[j] = fmaf(a.y, b.y, ex[j]); ex[j] = fmaf(a.z, b.z, ex[j]); ex[j] = fmaf(a.w, b.w, ex[j]);
            }
        }
    }
    float *p = red + (warp & 3) * RBv * RL;
    __syncthreads();   // readers of the previous pass' sums are done
    if (warp >= 4) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            *reinterpret_cast<float2 *>(p + g * RL + 8 * nt + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
            *reinterpret_cast<float2 *>(p + (g + 8) * RL + 8 * nt + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
        }
        if (RBv > 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[(16 + (lane >> 4)) * RL + (lane & 15) + 16 * j] = ex[j];
        }
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            float2 *q0 = reinterpret_cast<float2 *>(p + g * RL + 8 * nt + 2 * t), *q1 = reinterpret_cast<float2 *>(p + (g + 8) * RL + 8 * nt + 2 * t);
            const float2 u0 = *q0, u1 = *q1;
            *q0 = make_float2(u0.x + acc[nt][0], u0.y + acc[nt][1]);
            *q1 = make_float2(u1.x + acc[nt][2], u1.y + acc[nt][3]);
        }
        if (RBv > 16) {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[(16 + (lane >> 4)) * RL + (lane & 15) + 16 * j] += ex[j];
        }
    }
    __syncthreads();
}

template <int RI>
__device__ __forceinline__ float fwd_value(const float *red, int m, int n) {
    constexpr int RL = RedLd<RI>::v, S = 2 * RI * RL;
    const float *q = red + m * RL + n;
    return (q[0] + q[S]) + (q[2 * S] + q[3 * S]);
}

// ---- backward pass: partial dH1[m][k] = sum_{n in slice} dz[m][n] W[n][k]  (RB x 256, reduction over 64) -------
// dzs [64 n][2 mg][RIP] (a lane's RI rows mg + 2 i are contiguous; RIP = RI rounded up to 4), Ws [64][WLD],
// part [RB][256].  Warp w owns k in [64 (w&3), +64) and the n half (w>>2); the two halves are added through `part`.
// The 8 partial results of a cluster are summed by reduce_scatter_rows.  All 256 threads; contains __syncthreads.
template <int RI>
__device__ __forceinline__ void bwd_pass(const float *dzs, const float *Ws, float *part) {
    constexpr int RIP = (RI + 3) & ~3;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, mg = lane >> 4, kg = lane & 15, kq = warp & 3, nh = warp >> 2;
    const float *ap = dzs + (32 * nh) * 2 * RIP + mg * RIP;
    const float *bp = Ws + (32 * nh) * WLD + 64 * kq + 4 * kg;
    float acc[RI][4];
#pragma unroll
    for (int i = 0; i < RI; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
    for (int n = 0; n < 32; ++n) {
        float av[RIP];
#pragma unroll
        for (int u = 0; u < RIP / 4; ++u) {
            const float4 t = *reinterpret_cast<const float4 *>(ap + n * 2 * RIP + 4 * u);
            av[4 * u] = t.x; av[4 * u + 1] = t.y; av[4 * u + 2] = t.z; av[4 * u + 3] = t.w;
        }
        const float4 b = *reinterpret_cast<const float4 *>(bp + n * WLD);
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
            acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
            acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
            acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
        }
    }
    float *p = part + mg * H1 + 64 * kq + 4 * kg;
    if (nh == 1) {
#pragma unroll
        for (int i = 0; i < RI; ++i) *reinterpret_cast<float4 *>(p + 2 * i * H1) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    if (nh == 0) {
#pragma unroll
        for (int i = 0; i < RI; ++i) {
            float4 *q = reinterpret_cast<float4 *>(p + 2 * i * H1);
            const float4 t = *q;
            *q = make_float4(acc[i][0] + t.x, acc[i][1] + t.y, acc[i][2] + t.z, acc[i][3] + t.w);
        }
    }
    __syncthreads();
}

// After a cluster barrier: CTA `crank` collects rows crank + 8 q (its own first-layer rows) of the dH1 partials of all 8
// CTAs in a fixed order.  out[RPW][256] in local shared memory; thread t sums float4 (t & 63) of row slot (t >> 6).
template <int RB_>
__device__ __forceinline__ void reduce_scatter_rows(const float *part, int crank, float *out) {
    const int tid = threadIdx.x, q = tid >> 6, c4 = tid & 63, m = crank + 8 * q;
    if (m >= RB_) return;
    const uint32_t local = smem_u32(part + m * H1 + 4 * c4);
    float4 v[CL];
#pragma unroll
    for (int p = 0; p < CL; ++p) v[p] = ld_cluster_f4(mapa(local, p));
    float4 s = v[0];
#pragma unroll
    for (int p = 1; p < CL; ++p) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
    *reinterpret_cast<float4 *>(out + q * H1 + 4 * c4) = s;
}

// ---- dW2 tile: G[n][k] = sum_b dz2[b][n] h1[b][k0 + k]   (64 x TK outputs, reduction over the batch) ------------
// dz2 comes from the slice-major scratch DZ2S[..][rows][64] (a tile's operand is one contiguous block), h1 from the
// padded activation rows H1P[b][WLD].  Operands are staged with cp.async (16-byte chunks) into Dz[rows][64] and
// Hs[rows][TK]; the batch is split over the 8 warps; a lane owns n in {4 ng..+3, 32 + 4 ng..+3} and TK/4 consecutive k.
__device__ __forceinline__ void cp16(float *dst, const float *src, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    const int n = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// Stage rows b0 .. b0 + rows_pad - 1 (rows >= B are zero-filled; rows_pad multiple of 8).  dzs: this n-slice's block
// [Bpad][64]; h1p: padded activation rows.  All 256 threads issue; caller commits / waits.
template <int TK>
__device__ __forceinline__ void dw_stage(float *Dz, float *Hs, const float *dzs, const float *h1p, int k0, int b0, int B, int rows_pad) {
    for (int c = threadIdx.x; c < rows_pad * 16; c += CTHREADS) {
        const int r = c >> 4, q = c & 15, b = b0 + r;
        cp16(Dz + r * NS + 4 * q, dzs + (size_t)min(b, B - 1) * NS + 4 * q, b < B);
    }
    constexpr int CH = TK / 4;
    for (int c = threadIdx.x; c < rows_pad * CH; c += CTHREADS) {
        const int r = c / CH, q = c - r * CH, b = b0 + r;
        cp16(Hs + r * TK + 4 * q, h1p + (size_t)min(b, B - 1) * WLD + k0 + 4 * q, b < B);
    }
}
// accumulate the staged rows (optionally scaled): warp w takes rows w, w + 8, ...
template <int TK>
__device__ __forceinline__ void dw_accumulate(const float *Dz, const float *Hs, int rows_pad, float (&acc)[8][TK / 4]) {
    constexpr int KL = TK / 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, ng = lane & 7, kg = lane >> 3;
#pragma unroll 4
    for (int r = warp; r < rows_pad; r += 8) {
        float a[8], h[KL];
        const float4 a0 = *reinterpret_cast<const float4 *>(Dz + r * NS + 4 * ng), a1 = *reinterpret_cast<const float4 *>(Dz + r * NS + 32 + 4 * ng);
        a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
        for (int u = 0; u < KL / 4; ++u) {
            const float4 t = *reinterpret_cast<const float4 *>(Hs + r * TK + KL * kg + 4 * u);
            h[4 * u] = t.x; h[4 * u + 1] = t.y; h[4 * u + 2] = t.z; h[4 * u + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < KL; ++k) acc[i][k] = fmaf(a[i], h[k], acc[i][k]);
    }
}
// fixed-order sum of the 8 warps' partial tiles through red[4][TK][DLD] (transposed: n contiguous, DLD = 68).  On return
// thread t holds G[n][k] for k = t % TK, n = t / TK + (256 / TK) j, j < 64 TK / 256: lanes run along k, so the Adam
// stage's accesses to W2[n][k0 + k] are contiguous.  Contains barriers.
constexpr int DLD = NS + 4;
template <int TK>
__device__ __forceinline__ void dw_reduce(float (&acc)[8][TK / 4], float *red, float (&g)[TK / 4]) {
    constexpr int KL = TK / 4, NPT = TK / 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, ng = lane & 7, kg = lane >> 3;
    float *p = red + (warp & 3) * TK * DLD;
    __syncthreads();
    if (warp >= 4) {
#pragma unroll
        for (int k = 0; k < KL; ++k)
#pragma unroll
            for (int u = 0; u < 2; ++u)
                *reinterpret_cast<float4 *>(p + (KL * kg + k) * DLD + 32 * u + 4 * ng) = make_float4(acc[4 * u][k], acc[4 * u + 1][k], acc[4 * u + 2][k], acc[4 * u + 3][k]);
    }
    __syncthreads();
    if (warp < 4) {
#pragma unroll
        for (int k = 0; k < KL; ++k)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                float4 *q = reinterpret_cast<float4 *>(p + (KL * kg + k) * DLD + 32 * u + 4 * ng);
                const float4 t = *q;
                *q = make_float4(t.x + acc[4 * u][k], t.y + acc[4 * u + 1][k], t.z + acc[4 * u + 2][k], t.w + acc[4 * u + 3][k]);
            }
    }
    __syncthreads();
    const int kk = tid % TK, nn = tid / TK;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
        const float *q = red + kk * DLD + nn + (CTHREADS / TK) * j;
        g[j] = (q[0] + q[TK * DLD]) + (q[2 * TK * DLD] + q[3 * TK * DLD]);
    }
}

}  // namespace cl
}  // namespace hirl
