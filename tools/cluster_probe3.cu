// Probe 3: LDS.128 cost by access pattern, and forward-pass micro-tile variants (all 8 warps of one CTA per SM).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/cluster_gemm.cuh"
using namespace hirl;
using namespace hirl::cl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
constexpr int NT = 12;

template <int PAT>
__device__ __forceinline__ long long lds_test(const float *smem, float &sink) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int off;
    if (PAT == 0) off = 0;                              // all lanes same address
    else if (PAT == 1) off = (lane & 7) * 4;            // 8 distinct consecutive float4, 4 quarters identical
    else if (PAT == 2) off = lane * 4;                  // all distinct, consecutive
    else if (PAT == 3) off = (lane & 7) * WLD;          // 8 rows (stride 260), quarters identical  [B operand of fwd_pass]
    else off = (lane >> 3) * WLD;                       // 4 rows, quarter-wide broadcast [A operand of fwd_pass]
    const float *p = smem + off + warp * 4;
    float4 acc = make_float4(0, 0, 0, 0);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 64; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const float4 v = *reinterpret_cast<const float4 *>(p + ((it * 16 + u) & 31) * 8);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    sink += acc.x + acc.y + acc.z + acc.w;
    return (t1 - t0);
}

// V1: half-warp k-groups (16 groups x 16 k), lane tile 8 rows x 8 cols
__device__ __forceinline__ void fwd_pass_v1(const float *As, const float *Ws, float *red, float (&z)[4]) {
    const int tid = threadIdx.x, kgp = tid >> 4, l = tid & 15, mg = l >> 3, ng = l & 7;
    const float *ap = As + mg * WLD + 16 * kgp;   // rows mg + 2 i
    const float *bp = Ws + ng * WLD + 16 * kgp;   // rows ng + 8 j
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float4 a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4 *>(ap + 2 * i * WLD + 4 * r);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4 *>(bp + 8 * j * WLD + 4 * r);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
    // combine the two half-warps of a warp, then the 8 warps through smem (simple version for timing)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
    const int warp = tid >> 5;
    float *p = red + (warp & 3) * RB * RLD + mg * RLD + ng;
    if (warp >= 4 && l == (tid & 31)) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) p[2 * i * RLD + 8 * j] = acc[i][j];
    }
    __syncthreads();
    if (warp < 4 && l == (tid & 31)) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) p[2 * i * RLD + 8 * j] += acc[i][j];
    }
    __syncthreads();
    const float *q = red + (tid >> 4) * RLD + (tid & 15);
#pragma unroll
    for (int c = 0; c < 4; ++c) z[c] = (q[16 * c] + q[RB * RLD + 16 * c]) + (q[2 * RB * RLD + 16 * c] + q[3 * RB * RLD + 16 * c]);
}

// V5: FFMA only (operands loaded once): the issue-rate ceiling of the V2 loop body
__device__ __forceinline__ void fwd_pass_ffma_only(const float *As, const float *Ws, float (&z)[4]) {
    const int lane = threadIdx.x & 31, mg = lane >> 3, ng = lane & 7, warp = threadIdx.x >> 5;
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    float4 a[4], b[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4 *>(As + (mg + 4 * i) * WLD + 32 * warp);
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = *reinterpret_cast<const float4 *>(Ws + (ng + 8 * j) * WLD + 32 * warp);
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
            }
    }
    z[0] = z[1] = z[2] = z[3] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) z[(i + j) & 3] += acc[i][j];
}

__global__ void __launch_bounds__(CTHREADS, 1) probe3(long long *cycg, float *sinkg) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    float *Ab = smem + 2 * W_FLOATS, *red = Ab + A_FLOATS;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2 * W_FLOATS + A_FLOATS; i += CTHREADS) smem[i] = (float)(i & 1023) * 1e-3f;
    __syncthreads();
    long long *cyc = cycg + (size_t)blockIdx.x * NT;
    float s = 0.f, z[4];
    long long t;
    t = lds_test<0>(smem, s); if (tid == 0) cyc[0] = t;
    t = lds_test<1>(smem, s); if (tid == 0) cyc[1] = t;
    t = lds_test<2>(smem, s); if (tid == 0) cyc[2] = t;
    t = lds_test<3>(smem, s); if (tid == 0) cyc[3] = t;
    t = lds_test<4>(smem, s); if (tid == 0) cyc[4] = t;
    const int R = 16;
    __syncthreads();
    long long t0 = clock64();
    for (int r = 0; r < R; ++r) { fwd_pass(Ab, smem + (r & 1) * W_FLOATS, red, z); s += z[0] + z[3]; __syncthreads(); }
    if (tid == 0) cyc[5] = (clock64() - t0) / R;
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < R; ++r) { fwd_pass_v1(Ab, smem + (r & 1) * W_FLOATS, red, z); s += z[0] + z[3]; __syncthreads(); }
    if (tid == 0) cyc[6] = (clock64() - t0) / R;
    __syncthreads();
    t0 = clock64();
    for (int r = 0; r < R; ++r) { fwd_pass_ffma_only(Ab, smem + (r & 1) * W_FLOATS, z); s += z[0] + z[3]; __syncthreads(); }
    if (tid == 0) cyc[7] = (clock64() - t0) / R;
    if (s == 123.456f) sinkg[0] = s;
}

int main() {
    float *dsink; long long *dcyc;
    CK(cudaMalloc(&dsink, 4)); CK(cudaMalloc(&dcyc, 148 * NT * 8)); CK(cudaMemset(dcyc, 0, 148 * NT * 8));
    const size_t sm = (size_t)(2 * W_FLOATS + A_FLOATS + RED_FLOATS) * 4;
    CK(cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    const int grid = 148;
    for (int it = 0; it < 2; ++it) { probe3<<<grid, CTHREADS, sm>>>(dcyc, dsink); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
    std::vector<long long> cyc((size_t)grid * NT);
    CK(cudaMemcpy(cyc.data(), dcyc, cyc.size() * 8, cudaMemcpyDeviceToHost));
    const char *names[8] = {"LDS.128 x1024/warp, all lanes same address", "LDS.128 8 distinct consecutive (quarters identical)", "LDS.128 32 distinct consecutive",
                            "LDS.128 8 rows stride 260 (B operand pattern)", "LDS.128 4 rows quarter broadcast (A operand pattern)",
                            "fwd_pass V2 (4x8 lane tile, warp k-split)", "fwd_pass V1 (8x8 lane tile, half-warp k-groups)", "V2 FFMA only (no LDS in loop, no reduction)"};
    for (int t = 0; t < 8; ++t) {
        long long mn = 1LL << 60, mx = 0; double s = 0;
        for (int g = 0; g < grid; ++g) { const long long v = cyc[(size_t)g * NT + t]; mn = v < mn ? v : mn; mx = v > mx ? v : mx; s += v; }
        if (t < 5) printf("  %-55s cycles mean %9.1f  -> %.2f cycles per warp-LDS.128 (8 warps x 1024)\n", names[t], s / grid, (s / grid) / (8 * 1024.0));
        else printf("  %-55s cycles min %7lld mean %9.1f max %7lld\n", names[t], mn, s / grid, mx);
    }
    return 0;
}
