// Latency probe for design decisions (not part of the product): dependent-load round trips on B200.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>
#include <numeric>
#include <random>

template <int MODE>
__global__ void chase(const unsigned *buf, int iters, unsigned *out, long long *cycles) {
    unsigned i = 0;
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) {
        if (MODE == 0) i = buf[i];
        else if (MODE == 1) i = __ldcg(buf + i);
        else if (MODE == 2) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(buf + i) : "memory"); i = v; }
        else if (MODE == 3) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(buf + i) : "memory"); i = v; }
        else if (MODE == 4) i = atomicAdd((unsigned *)buf + i, 0u);
    }
    long long t1 = clock64();
    out[0] = i;
    cycles[0] = t1 - t0;
}
// fence cost
__global__ void fence_cost(int iters, long long *cycles, unsigned *sink) {
    long long t0 = clock64();
    for (int k = 0; k < iters; ++k) { sink[threadIdx.x] = k; __threadfence(); }
    cycles[0] = clock64() - t0;
}
// ping-pong between two CTAs on different SMs through global flags
__global__ void pingpong(unsigned *flags, int iters, long long *cycles) {
    const int me = blockIdx.x;
    long long t0 = clock64();
    for (int k = 1; k <= iters; ++k) {
        if (me == 0) {
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags), "r"((unsigned)k) : "memory");
            unsigned v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + 32) : "memory"); } while (v < (unsigned)k);
        } else {
            unsigned v;
            do { asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags) : "memory"); } while (v < (unsigned)k);
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flags + 32), "r"((unsigned)k) : "memory");
        }
    }
    if (me == 0) cycles[0] = clock64() - t0;
}
int main() {
    const int N = 1 << 20;  // 4 MB: L2 resident, larger than L1
    std::vector<unsigned> h(N);
    std::vector<unsigned> perm(N);
    std::iota(perm.begin(), perm.end(), 0u);
    std::mt19937 rng(1);
    std::shuffle(perm.begin(), perm.end(), rng);
    for (int i = 0; i < N; ++i) h[perm[i]] = perm[(i + 1) % N];
    unsigned *d, *out, *flags; long long *cyc;
    cudaMalloc(&d, N * 4); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&flags, 1024);
    cudaMemcpy(d, h.data(), N * 4, cudaMemcpyHostToDevice);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 20000;
    const char *names[] = {"ld (weak, L1 miss -> L2)", "ld.cg", "ld.relaxed.gpu", "ld.acquire.gpu", "atomicAdd(+0) round trip"};
    for (int rep = 0; rep < 2; ++rep) {
        for (int m = 0; m < 5; ++m) {
            if (m == 0) chase<0><<<1, 1>>>(d, iters, out, cyc);
            if (m == 1) chase<1><<<1, 1>>>(d, iters, out, cyc);
            if (m == 2) chase<2><<<1, 1>>>(d, iters, out, cyc);
            if (m == 3) chase<3><<<1, 1>>>(d, iters, out, cyc);
            if (m == 4) chase<4><<<1, 1>>>(d, iters, out, cyc);
            long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            if (rep) printf("%-28s %8.1f cycles  %7.1f ns (clockRate %d kHz)\n", names[m], (double)c / iters, (double)c / iters / (clk * 1e-6), clk);
        }
    }
    fence_cost<<<1, 32>>>(2000, cyc, out);
    { long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("store + __threadfence            %8.1f cycles\n", (double)c / 2000); }
    cudaMemset(flags, 0, 1024);
    pingpong<<<2, 1>>>(flags, 2000, cyc);
    { long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("2-CTA flag ping-pong round trip   %8.1f cycles  %7.1f ns\n", (double)c / 2000, (double)c / 2000 / (clk * 1e-6)); }
    // small working set (L1 hits for weak loads)
    chase<0><<<1, 1>>>(d, 64, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
