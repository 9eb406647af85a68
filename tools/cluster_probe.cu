// Stand-alone correctness + timing probe of csrc/cluster_gemm.cuh (the building blocks of the cluster-resident update):
// cluster barrier, DSMEM exchange round, hierarchical grid barrier, streamed forward pass, dH1 pass + reduce-scatter,
// dW2 tile.  128 CTAs = 16 clusters of 8, all timed with clock64 (min / mean / max over CTAs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/cluster_probe tools/cluster_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/cluster_gemm.cuh"
using namespace hirl;
using namespace hirl::cl;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int NT = 9;   // timed sections
struct Args {
    const float *W2a, *W2b, *H1g, *DZ2;
    float *Z, *DH1, *GW, *GW32;
    unsigned *flags;
    long long *cyc;   // [grid][NT]
    int B, reps;
};

__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned *p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

__device__ __forceinline__ void grid_sync(unsigned *flags, unsigned &gen, int nclusters, int cid, int crank) {
    cluster_sync_all();
    gen += 1u;
    if (crank == 0 && threadIdx.x < 32) {
        if (threadIdx.x == 0) st_release_u32(flags + cid * 32, gen);
        bool pending = true;
        const long long t0 = clock64();
        while (pending) {
            pending = (int)threadIdx.x < nclusters && (int)(ld_relaxed_u32(flags + threadIdx.x * 32) - gen) < 0;
            pending = __any_sync(0xffffffffu, pending);
            if (clock64() - t0 > 4000000000LL) { printf("grid_sync timeout block %d\n", blockIdx.x); __trap(); }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    cluster_sync_all();
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CTHREADS, 1) probe_kernel(const Args a) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
        float *Ab = smem + 2 * W_FLOATS;
    float *red = Ab + A_FLOATS;
    float *dzT = red + RED_FLOATS;          // [64][16]
    float *rows2 = dzT + NS * RB;           // [2][256]
    float *xb = rows2 + 2 * H1;             // [2][8][64] exchange
    __shared__ uint64_t wbar[2], abar;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)cluster_ctarank(), cid = (int)cluster_id_x(), ncl = gridDim.x / CL;
    if (tid == 0) { mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); mbar_init(&abar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    long long *cyc = a.cyc + (size_t)blockIdx.x * NT;
    long long t0;
    const int b0 = cid * RB;

    // ---- T0: cluster barrier
    cluster_sync_all();
    t0 = clock64();
    for (int i = 0; i < 64; ++i) cluster_sync_all();
    if (tid == 0) cyc[0] = (clock64() - t0) / 64;

    // ---- T1: DSMEM exchange round: every half-warp posts one value to all peers, barrier, read 8
    float accx = 0.f;
    t0 = clock64();
    for (int i = 0; i < 32; ++i) {
        float *buf = xb + (i & 1) * CL * 64;
        const int hw = tid >> 4, l16 = tid & 15;
        if (l16 < CL) st_cluster_f32(mapa(smem_u32(buf + crank * 64 + hw), l16), (float)(i + hw + crank));
        cluster_sync_all();
        float s = 0.f;
#pragma unroll
        for (int p = 0; p < CL; ++p) s += buf[p * 64 + hw];
        accx += s;
    }
    if (tid == 0) cyc[1] = (clock64() - t0) / 32;

    // ---- T8: the same exchange with st.async + the receiver's mbarrier (no cluster barrier)
    {
        __shared__ uint64_t xbar[2];
        if (tid == 0) { mbar_init(&xbar[0], 1); mbar_init(&xbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        cluster_sync_all();
        uint32_t xph[2] = {0u, 0u};
        t0 = clock64();
        for (int i = 0; i < 32; ++i) {
            const int par = i & 1;
            float *buf = xb + par * CL * 64;
            const int hw = tid >> 4, l16 = tid & 15;
            if (tid == 0) mbar_expect_tx(&xbar[par], CL * 16 * 4);
            if (l16 < CL) st_async_f32(mapa(smem_u32(buf + crank * 64 + hw), l16), (float)(i + hw + crank), mapa(smem_u32(&xbar[par]), l16));
            mbar_wait(&xbar[par], xph[par]);
            xph[par] ^= 1u;
            float s = 0.f;
#pragma unroll
            for (int p = 0; p < CL; ++p) s += buf[p * 64 + hw];
            accx += s;
        }
        if (tid == 0) cyc[8] = (clock64() - t0) / 32;
        cluster_sync_all();
    }

    // ---- T5: hierarchical grid barrier
    unsigned gen = ld_relaxed_u32(a.flags + cid * 32);
    grid_sync(a.flags, gen, ncl, cid, crank);
    t0 = clock64();
    for (int i = 0; i < 32; ++i) grid_sync(a.flags, gen, ncl, cid, crank);
    if (tid == 0) cyc[5] = (clock64() - t0) / 32;

    // ---- T2: forward passes with double-buffered weight streaming
    uint32_t wph[2] = {0u, 0u};
    if (warp == 0) {
        load_wslice_warp(smem, a.W2a, NS * crank, &wbar[0], lane);
        load_ablock_warp(Ab, a.H1g, b0, a.B, &abar, lane);
    }
    mbar_wait(&abar, 0u);
    float z[4];
    long long t_first = 0;
    t0 = clock64();
    for (int rep = 0; rep < a.reps; ++rep) {
        const int cur = rep & 1;
        if (warp == 0 && rep + 1 < a.reps) load_wslice_warp((smem + (cur ^ 1) * W_FLOATS), ((rep + 1) & 1) ? a.W2b : a.W2a, NS * crank, &wbar[cur ^ 1], lane);
        mbar_wait(&wbar[cur], wph[cur]);
        wph[cur] ^= 1u;
        if (rep == 0) t_first = clock64() - t0;
        fwd_pass(Ab, (smem + cur * W_FLOATS), red, z);
        __syncthreads();
        if (rep == 0) t0 = clock64();   // steady state from the second pass on
    }
    if (tid == 0) { cyc[2] = (clock64() - t0) / (a.reps - 1); cyc[6] = t_first; }
    // the last pass used W2a when reps is odd
    {
        const int m = tid >> 4, c = tid & 15;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (b0 + m < a.B) a.Z[(size_t)(b0 + m) * H2 + NS * crank + c + 16 * q] = z[q] + accx * 0.f;
    }

    // ---- T3: dH1 pass + reduce-scatter (weights of the last pass are resident in Wb[(reps-1)&1])
    const float *Wlast = smem + ((a.reps - 1) & 1) * W_FLOATS;
    for (int e = tid; e < NS * RB; e += CTHREADS) {
        const int n = e >> 4, m = e & 15;
        dzT[e] = a.DZ2[(size_t)min(b0 + m, a.B - 1) * H2 + NS * crank + n];
    }
    __syncthreads();
    t0 = clock64();
    for (int rep = 0; rep < 8; ++rep) {
        bwd_pass(dzT, Wlast, red);
        cluster_sync_all();
        reduce_scatter_rows(red, crank, rows2);
        cluster_sync_all();
    }
    if (tid == 0) cyc[3] = (clock64() - t0) / 8;
    for (int e = tid; e < 2 * H1; e += CTHREADS) {
        const int b = b0 + 2 * crank + (e >> 8);
        if (b < a.B) a.DH1[(size_t)b * H1 + (e & 255)] = rows2[e];
    }
    __syncthreads();

    // ---- T4: dW2 tiles (64 x 32 and 32 x 32 outputs, reduction over the batch)
    {
        float *Dz = smem, *Hs = smem + 256 * 64, *rd = Hs + 256 * 32;
        const int tile = blockIdx.x & 63, n0 = (tile >> 3) * 64, k0 = (tile & 7) * 32;
        t0 = clock64();
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
        for (int c0 = 0; c0 < a.B; c0 += 256) {
            const int rows = min(256, a.B - c0), rp = (rows + 7) & ~7;
            __syncthreads();
            dw_stage<64>(Dz, Hs, a.DZ2, H2, n0, a.H1g, k0, c0, a.B, rp);
            cp_commit(); cp_wait_all();
            __syncthreads();
            dw_accumulate<64>(Dz, Hs, rp, acc);
        }
        float g[8];
        dw_reduce<64>(acc, rd, g);
        if (tid == 0) cyc[4] = clock64() - t0;
        const int n = tid & 63, kq = tid >> 6;
        if (blockIdx.x < 64)
#pragma unroll
            for (int j = 0; j < 8; ++j) a.GW[(size_t)(n0 + n) * H1 + k0 + 8 * kq + j] = g[j];
        __syncthreads();
    }
    {
        float *Dz = smem, *Hs = smem + 256 * 32, *rd = Hs + 256 * 32;
        const int tile = blockIdx.x, n0 = ((tile >> 3) & 15) * 32, k0 = (tile & 7) * 32;
        t0 = clock64();
        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;
        for (int c0 = 0; c0 < a.B; c0 += 256) {
            const int rows = min(256, a.B - c0), rp = (rows + 7) & ~7;
            __syncthreads();
            dw_stage<32>(Dz, Hs, a.DZ2, H2, n0, a.H1g, k0, c0, a.B, rp);
            cp_commit(); cp_wait_all();
            __syncthreads();
            dw_accumulate<32>(Dz, Hs, rp, acc);
        }
        float g[4];
        dw_reduce<32>(acc, rd, g);
        if (tid == 0) cyc[7] = clock64() - t0;
        const int n = tid & 31, kq = tid >> 5;
#pragma unroll
        for (int j = 0; j < 4; ++j) a.GW32[(size_t)(n0 + n) * H1 + k0 + 4 * kq + j] = g[j];
    }
}

static double relerr(const std::vector<double> &ref, const float *got, size_t n) {
    double num = 0, den = 0;
    for (size_t i = 0; i < n; ++i) { const double d = got[i] - ref[i]; num += d * d; den += ref[i] * ref[i]; }
    return std::sqrt(num / (den + 1e-300));
}

int main(int argc, char **argv) {
    const int B = argc > 1 ? atoi(argv[1]) : 256, reps = 17;
    int grid = 128;
    std::vector<float> hWa((size_t)H2 * H1), hWb((size_t)H2 * H1), hH((size_t)B * H1), hD((size_t)B * H2);
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
    for (auto &v : hWa) v = rnd();
    for (auto &v : hWb) v = rnd();
    for (auto &v : hH) v = rnd();
    for (auto &v : hD) v = rnd();
    float *dWa, *dWb, *dH, *dD, *dZ, *dDH1, *dGW, *dGW32;
    unsigned *dflags;
    long long *dcyc;
    CK(cudaMalloc(&dWa, hWa.size() * 4)); CK(cudaMalloc(&dWb, hWb.size() * 4)); CK(cudaMalloc(&dH, hH.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4));
    CK(cudaMalloc(&dZ, (size_t)B * H2 * 4)); CK(cudaMalloc(&dDH1, (size_t)B * H1 * 4)); CK(cudaMalloc(&dGW, (size_t)H2 * H1 * 4)); CK(cudaMalloc(&dGW32, (size_t)H2 * H1 * 4));
    CK(cudaMalloc(&dflags, 64 * 32 * 4)); CK(cudaMemset(dflags, 0, 64 * 32 * 4));
    CK(cudaMalloc(&dcyc, (size_t)128 * NT * 8)); CK(cudaMemset(dcyc, 0, (size_t)128 * NT * 8));
    CK(cudaMemcpy(dWa, hWa.data(), hWa.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dWb, hWb.data(), hWb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dH, hH.data(), hH.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dD, hD.data(), hD.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dZ, 0, (size_t)B * H2 * 4)); CK(cudaMemset(dDH1, 0, (size_t)B * H1 * 4));
    const size_t smem = (size_t)(2 * W_FLOATS + A_FLOATS + RED_FLOATS + NS * RB + 2 * H1 + 2 * CL * 64) * 4;
    const size_t smem2 = (size_t)(256 * 64 + 256 * 32 + 4 * 32 * 64) * 4;
    const size_t sm = smem > smem2 ? smem : smem2;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(CTHREADS); cfg.dynamicSmemBytes = sm;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int ncl = 0;
        CK(cudaOccupancyMaxActiveClusters(&ncl, probe_kernel, &cfg));
        printf("smem %zu B, max active clusters of %d: %d\n", sm, CL, ncl);
        if (ncl * CL < grid) grid = ncl * CL;
    }
    Args a{dWa, dWb, dH, dD, dZ, dDH1, dGW, dGW32, dflags, dcyc, B, reps};
    for (int it = 0; it < 2; ++it) {
        probe_kernel<<<grid, CTHREADS, sm>>>(a);
        CK(cudaGetLastError());
        CK(cudaDeviceSynchronize());
    }
    std::vector<long long> cyc((size_t)grid * NT);
    CK(cudaMemcpy(cyc.data(), dcyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
    const char *names[NT] = {"cluster barrier", "DSMEM exchange round (post + barrier + read)", "fwd pass 16x64x256 incl. streamed slice (steady)",
                             "dH1 pass + 2 cluster barriers + reduce-scatter", "dW2 tile 64x32 over the batch (stage+acc+reduce)", "grid barrier (hierarchical)",
                             "first pass: cold wait for slice + A block", "dW2 tile 32x32", "DSMEM exchange round via st.async + mbarrier"};
    for (int t = 0; t < NT; ++t) {
        long long mn = 1LL << 60, mx = 0; double s = 0;
        for (int g = 0; g < grid; ++g) { const long long v = cyc[(size_t)g * NT + t]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; s += v; }
        printf("  %-58s cycles min %7lld mean %9.1f max %7lld\n", names[t], mn, s / grid, mx);
    }
    // verification
    std::vector<float> gZ((size_t)B * H2), gDH((size_t)B * H1), gGW((size_t)H2 * H1), gGW32((size_t)H2 * H1);
    CK(cudaMemcpy(gZ.data(), dZ, gZ.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gDH.data(), dDH1, gDH.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gGW.data(), dGW, gGW.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gGW32.data(), dGW32, gGW32.size() * 4, cudaMemcpyDeviceToHost));
    const std::vector<float> &Wl = ((reps - 1) & 1) ? hWb : hWa;
    std::vector<double> rZ((size_t)B * H2), rDH((size_t)B * H1), rGW((size_t)H2 * H1);
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < H2; ++n) { double s = 0; for (int k = 0; k < H1; ++k) s += (double)hH[(size_t)b * H1 + k] * Wl[(size_t)n * H1 + k]; rZ[(size_t)b * H2 + n] = s; }
    for (int b = 0; b < B; ++b)
        for (int k = 0; k < H1; ++k) { double s = 0; for (int n = 0; n < H2; ++n) s += (double)hD[(size_t)b * H2 + n] * Wl[(size_t)n * H1 + k]; rDH[(size_t)b * H1 + k] = s; }
    for (int n = 0; n < H2; ++n)
        for (int k = 0; k < H1; ++k) { double s = 0; for (int b = 0; b < B; ++b) s += (double)hD[(size_t)b * H2 + n] * hH[(size_t)b * H1 + k]; rGW[(size_t)n * H1 + k] = s; }
    const int nrows = std::min(B, (grid / CL) * RB);
    printf("rel-L2 vs fp64: fwd %.3e  dH1 %.3e  dW2(64) %.3e  dW2(32) %.3e  (rows checked %d)\n", relerr(rZ, gZ.data(), (size_t)nrows * H2),
           relerr(rDH, gDH.data(), (size_t)nrows * H1), relerr(rGW, gGW.data(), rGW.size()), relerr(rGW, gGW32.data(), rGW.size()), nrows);
    return 0;
}
