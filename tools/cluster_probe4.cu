// Probe 4: forward pass on mma.sync (3xTF32) vs the FFMA version: cycles per pass (resident operands) and accuracy vs fp64.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/cluster_gemm.cuh"
using namespace hirl;
using namespace hirl::cl;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <int RI, int MODE>
__global__ void __launch_bounds__(CTHREADS, 1) probe4(const float *A, const float *W, float *Z, long long *cyc) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    constexpr int RB = 2 * RI, RL = RedLd<RI>::v;
    float *Ws = smem, *As = smem + W_FLOATS, *red = As + RB * WLD;
    const int tid = threadIdx.x;
    for (int i = tid; i < NS * H1; i += CTHREADS) Ws[(i >> 8) * WLD + (i & 255)] = W[(size_t)(blockIdx.x % 8) * NS * H1 + i];
    for (int i = tid; i < RB * H1; i += CTHREADS) As[(i >> 8) * WLD + (i & 255)] = A[i];
    __syncthreads();
    float s = 0.f;
    long long t0 = clock64();
    for (int r = 0; r < 16; ++r) {
        if (MODE == 0) fwd_pass<RI>(As, Ws, red); else fwd_pass_mma<RI>(As, Ws, red);
        s += fwd_value<RI>(red, tid >> 4, tid & 15);
        __syncthreads();
    }
    if (tid == 0) cyc[blockIdx.x] = (clock64() - t0) / 16;
    if (blockIdx.x < 8)
        for (int i = tid; i < RB * NS; i += CTHREADS) Z[(size_t)(i >> 6) * H2 + NS * blockIdx.x + (i & 63)] = fwd_value<RI>(red, i >> 6, i & 63) + s * 0.f;
}

template <int RI, int MODE>
static void run(const char *name, const float *dA, const float *dW, float *dZ, long long *dcyc, const std::vector<float> &hA, const std::vector<float> &hW) {
    constexpr int RB = 2 * RI;
    const size_t sm = (size_t)(W_FLOATS + RB * WLD + 4 * RB * 80) * 4;
    CK(cudaFuncSetAttribute(probe4<RI, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    for (int it = 0; it < 2; ++it) { probe4<RI, MODE><<<148, CTHREADS, sm>>>(dA, dW, dZ, dcyc); CK(cudaGetLastError()); CK(cudaDeviceSynchronize()); }
    std::vector<long long> cyc(148);
    std::vector<float> z((size_t)RB * H2);
    CK(cudaMemcpy(cyc.data(), dcyc, 148 * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(z.data(), dZ, z.size() * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0, mx = 0;
    for (int m = 0; m < RB; ++m)
        for (int n = 0; n < H2; ++n) {
            double ref = 0;
            for (int k = 0; k < H1; ++k) ref += (double)hA[(size_t)m * H1 + k] * hW[(size_t)n * H1 + k];
            const double d = z[(size_t)m * H2 + n] - ref;
            num += d * d; den += ref * ref; mx = std::fmax(mx, std::fabs(d));
        }
    double s = 0; for (auto v : cyc) s += v;
    printf("  %-34s cycles/pass mean %8.1f   rel-L2 vs fp64 %.3e  max abs %.3e\n", name, s / 148, std::sqrt(num / den), mx);
}

int main() {
    std::vector<float> hA((size_t)18 * H1), hW((size_t)H2 * H1);
    srand(3);
    for (auto &v : hA) v = (float)rand() / RAND_MAX * 2.f;          // post-ReLU-like (non-negative) activations
    for (auto &v : hW) v = (float)rand() / RAND_MAX - 0.5f;
    float *dA, *dW, *dZ; long long *dcyc;
    CK(cudaMalloc(&dA, hA.size() * 4)); CK(cudaMalloc(&dW, hW.size() * 4)); CK(cudaMalloc(&dZ, (size_t)18 * H2 * 4)); CK(cudaMalloc(&dcyc, 148 * 8));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dW, hW.data(), hW.size() * 4, cudaMemcpyHostToDevice));
    run<8, 0>("FFMA  RB=16", dA, dW, dZ, dcyc, hA, hW);
    run<8, 1>("mma.sync 3xTF32 RB=16", dA, dW, dZ, dcyc, hA, hW);
    run<9, 0>("FFMA  RB=18", dA, dW, dZ, dcyc, hA, hW);
    run<9, 1>("mma.sync 3xTF32 + FFMA rows RB=18", dA, dW, dZ, dcyc, hA, hW);
    return 0;
}
