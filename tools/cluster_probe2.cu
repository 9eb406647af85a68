// Probe 2: where do the cycles of a streamed forward pass go?  (a) compute only, (b) weight streaming only with several
// issue orders / copy flavours, on all co-resident clusters at once.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/cluster_gemm.cuh"
using namespace hirl;
using namespace hirl::cl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int NT = 10;
struct Args { const float *W; float *sink; long long *cyc; int nets; };

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(CTHREADS, 1) probe2(const Args a) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    float *Ab = smem + 2 * W_FLOATS, *red = Ab + A_FLOATS;
    __shared__ uint64_t wbar[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int crank = (int)cluster_ctarank(), cid = (int)cluster_id_x();
    if (tid == 0) { mbar_init(&wbar[0], 1); mbar_init(&wbar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    for (int i = tid; i < 2 * W_FLOATS + A_FLOATS; i += CTHREADS) smem[i] = (float)(i & 1023) * 1e-3f;
    __syncthreads();
    long long *cyc = a.cyc + (size_t)blockIdx.x * NT;
    long long t0;
    uint32_t ph[2] = {0u, 0u};
    float z[4], s = 0.f;
    const int R = 16;
    // (0) compute only
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) { fwd_pass(Ab, smem + (r & 1) * W_FLOATS, red, z); s += z[0] + z[3]; __syncthreads(); }
    if (tid == 0) cyc[0] = (clock64() - t0) / R;
    // (1) streaming only, bulk copies of 1 KB rows, same order everywhere, one slice in flight
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const float *w = a.W + (size_t)(r % a.nets) * H2 * H1;
        if (warp == 0) load_wslice_warp(smem + (r & 1) * W_FLOATS, w, NS * crank, &wbar[r & 1], lane);
        mbar_wait(&wbar[r & 1], ph[r & 1]); ph[r & 1] ^= 1u;
        __syncthreads();
    }
    if (tid == 0) cyc[1] = (clock64() - t0) / R;
    // (2) two slices in flight
    cluster_sync_all();
    t0 = clock64();
    if (warp == 0) load_wslice_warp(smem, a.W, NS * crank, &wbar[0], lane);
    for (int r = 0; r < R; ++r) {
        if (warp == 0 && r + 1 < R) load_wslice_warp(smem + ((r + 1) & 1) * W_FLOATS, a.W + (size_t)((r + 1) % a.nets) * H2 * H1, NS * crank, &wbar[(r + 1) & 1], lane);
        mbar_wait(&wbar[r & 1], ph[r & 1]); ph[r & 1] ^= 1u;
        __syncthreads();
    }
    if (tid == 0) cyc[2] = (clock64() - t0) / R;
    // (3) staggered row order per cluster, one slice in flight
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const float *w = a.W + (size_t)(r % a.nets) * H2 * H1;
        float *dst = smem + (r & 1) * W_FLOATS;
        if (warp == 0) {
            if (lane == 0) mbar_expect_tx(&wbar[r & 1], NS * H1 * 4);
            __syncwarp();
            for (int i = 0; i < 2; ++i) {
                const int n = (lane + 32 * i + 4 * cid) & 63;
                bulk_g2s(dst + n * WLD, w + (size_t)(NS * crank + n) * H1, H1 * 4, &wbar[r & 1]);
            }
        }
        mbar_wait(&wbar[r & 1], ph[r & 1]); ph[r & 1] ^= 1u;
        __syncthreads();
    }
    if (tid == 0) cyc[3] = (clock64() - t0) / R;
    // (4) cp.async 16-byte copies by all threads (LDGSTS), one slice in flight
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const float *w = a.W + (size_t)(r % a.nets) * H2 * H1 + (size_t)NS * crank * H1;
        float *dst = smem + (r & 1) * W_FLOATS;
        for (int c = tid; c < NS * 64; c += CTHREADS) { const int n = c >> 6, q = c & 63; cp16(dst + n * WLD + 4 * q, w + (size_t)n * H1 + 4 * q, true); }
        cp_commit(); cp_wait_all();
        __syncthreads();
    }
    if (tid == 0) cyc[4] = (clock64() - t0) / R;
    // (5) plain coalesced LDG.128 into registers (no smem), whole slice
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const float4 *w = reinterpret_cast<const float4 *>(a.W + (size_t)(r % a.nets) * H2 * H1 + (size_t)NS * crank * H1);
        float4 v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __ldcg(w + tid + 256 * i);
#pragma unroll
        for (int i = 0; i < 16; ++i) s += v[i].x + v[i].w;
        __syncthreads();
    }
    if (tid == 0) cyc[5] = (clock64() - t0) / R;
    // (6) one bulk copy of the whole contiguous 64 KB slice (unpadded destination)
    cluster_sync_all();
    t0 = clock64();
    for (int r = 0; r < R; ++r) {
        const float *w = a.W + (size_t)(r % a.nets) * H2 * H1 + (size_t)NS * crank * H1;
        if (tid == 0) { mbar_expect_tx(&wbar[r & 1], NS * H1 * 4); bulk_g2s(smem + (r & 1) * W_FLOATS, w, NS * H1 * 4, &wbar[r & 1]); }
        mbar_wait(&wbar[r & 1], ph[r & 1]); ph[r & 1] ^= 1u;
        __syncthreads();
    }
    if (tid == 0) cyc[6] = (clock64() - t0) / R;
    // (7) compute + prefetch of the next slice (the real loop), staggered order
    cluster_sync_all();
    t0 = clock64();
    if (warp == 0) load_wslice_warp(smem, a.W, NS * crank, &wbar[0], lane);
    for (int r = 0; r < R; ++r) {
        if (warp == 0 && r + 1 < R) {
            float *dst = smem + ((r + 1) & 1) * W_FLOATS;
            const float *w = a.W + (size_t)((r + 1) % a.nets) * H2 * H1;
            if (lane == 0) mbar_expect_tx(&wbar[(r + 1) & 1], NS * H1 * 4);
            __syncwarp();
            for (int i = 0; i < 2; ++i) {
                const int n = (lane + 32 * i + 4 * cid) & 63;
                bulk_g2s(dst + n * WLD, w + (size_t)(NS * crank + n) * H1, H1 * 4, &wbar[(r + 1) & 1]);
            }
        }
        mbar_wait(&wbar[r & 1], ph[r & 1]); ph[r & 1] ^= 1u;
        fwd_pass(Ab, smem + (r & 1) * W_FLOATS, red, z); s += z[1];
        __syncthreads();
    }
    if (tid == 0) cyc[7] = (clock64() - t0) / R;
    if (s == 123.456f) a.sink[0] = s;
}

int main(int argc, char **argv) {
    const int nets = argc > 1 ? atoi(argv[1]) : 7;
    float *dW, *dsink; long long *dcyc;
    CK(cudaMalloc(&dW, (size_t)nets * H2 * H1 * 4)); CK(cudaMemset(dW, 0, (size_t)nets * H2 * H1 * 4));
    CK(cudaMalloc(&dsink, 4)); CK(cudaMalloc(&dcyc, 128 * NT * 8)); CK(cudaMemset(dcyc, 0, 128 * NT * 8));
    const size_t sm = (size_t)(2 * W_FLOATS + A_FLOATS + RED_FLOATS) * 4;
    CK(cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(128); cfg.blockDim = dim3(CTHREADS); cfg.dynamicSmemBytes = sm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    CK(cudaOccupancyMaxActiveClusters(&ncl, probe2, &cfg));
    for (int use = ncl; use >= 1; use = (use == ncl && ncl > 8) ? 8 : (use > 1 ? 1 : 0)) {
        const int grid = use * CL;
        Args a{dW, dsink, dcyc, nets};
        for (int it = 0; it < 2; ++it) { probe2<<<grid, CTHREADS, sm>>>(a); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
        std::vector<long long> cyc((size_t)grid * NT);
        CK(cudaMemcpy(cyc.data(), dcyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
        const char *names[8] = {"compute only (fwd_pass, resident operands)", "stream: bulk 1KB rows, 1 in flight", "stream: bulk rows, 2 in flight",
                                "stream: bulk rows staggered per cluster", "stream: cp.async 16B all threads", "stream: LDG.128 to registers",
                                "stream: ONE 64KB bulk copy", "compute + prefetch (real loop, staggered)"};
        printf("clusters %d (grid %d), %d nets\n", use, grid, nets);
        for (int t = 0; t < 8; ++t) {
            long long mn = 1LL << 60, mx = 0; double s = 0;
            for (int g = 0; g < grid; ++g) { const long long v = cyc[(size_t)g * NT + t]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; s += v; }
            printf("  %-50s cycles min %7lld mean %9.1f max %7lld  (%.1f B/clk)\n", names[t], mn, s / grid, mx, 65536.0 / (s / grid));
        }
    }
    return 0;
}
