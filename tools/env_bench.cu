// env kernel variants: bandwidth + exhaustive exactness check of the constant-divisor division (not part of the product).
#include <cstdio>
#include <vector>
#include <cstring>
#include "../hirl4ucav_b200/csrc/env.cuh"
using namespace hirl;

__global__ void div_check(unsigned long long *bad) {
    const double D[3] = {10000.0, 3.141592653589793, 180.0};
    unsigned long long local = 0;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += (unsigned long long)gridDim.x * blockDim.x) {
        const float xf = __uint_as_float((unsigned)i);
        if (!(fabsf(xf) <= 3.0e38f)) continue;  // inf / nan
        const double x = (double)xf;
        for (int k = 0; k < 3; ++k) {
            const double d = D[k], r = 1.0 / d;
            const double q = div_const(x, d, r);
            const double ref = x / d;
            if (__double_as_longlong(q) != __double_as_longlong(ref)) { ++local; printf("mismatch: x=%.9g (0x%08x) d=%.17g fast=%.17g ieee=%.17g\n", xf, (unsigned)i, d, q, ref); }
        }
    }
    if (local) atomicAdd(bad, local);
}

int main() {
    const long long n = 10000000;
    std::vector<float> h((size_t)n * 16);
    unsigned s = 12345;
    for (size_t i = 0; i < h.size(); ++i) { s = s * 1664525u + 1013904223u; h[i] = (float)((int)(s >> 8) % 20000) * 0.5f - 3000.f; }
    for (long long i = 0; i < n; ++i) { unsigned f = (unsigned)(i * 2654435761u) & 31u; memcpy(&h[i * 16 + 15], &f, 4); h[i * 16 + 13] = (i % 7) * 0.05f; }
    float *raw, *obs, *rew; int *fl;
    cudaMalloc(&raw, n * 64); cudaMalloc(&obs, n * 52); cudaMalloc(&rew, n * 4); cudaMalloc(&fl, n * 4);
    cudaMemcpy(raw, h.data(), n * 64, cudaMemcpyHostToDevice);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mult : {4, 8, 16, 32}) {
        int grid = sms * mult;
        for (int w = 0; w < 3; ++w) env_step_kernel<<<grid, ENV_THREADS>>>((const float4 *)raw, n, obs, rew, fl, nullptr);
        cudaEventRecord(e0);
        for (int r = 0; r < 20; ++r) env_step_kernel<<<grid, ENV_THREADS>>>((const float4 *)raw, n, obs, rew, fl, nullptr);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("min_blocks %d fast_div %d grid %dxSM: %.1f us/pass  %.0f GB/s (124 B/state)\n", HIRL_ENV_MIN_BLOCKS, HIRL_ENV_FAST_DIV, mult, ms * 1e3 / 20, n * 124.0 / (ms / 20 * 1e-3) / 1e9);
    }
    // checksum of outputs (to compare variants bit for bit)
    std::vector<float> ho((size_t)n * 13), hr(n); std::vector<int> hf(n);
    cudaMemcpy(ho.data(), obs, n * 52, cudaMemcpyDeviceToHost); cudaMemcpy(hr.data(), rew, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(hf.data(), fl, n * 4, cudaMemcpyDeviceToHost);
    unsigned long long cs = 0;
    for (size_t i = 0; i < ho.size(); ++i) { unsigned u; memcpy(&u, &ho[i], 4); cs = cs * 1099511628211ull + u; }
    for (long long i = 0; i < n; ++i) { unsigned u; memcpy(&u, &hr[i], 4); cs = cs * 1099511628211ull + u + (unsigned)hf[i]; }
    printf("output checksum %016llx\n", cs);
#if HIRL_ENV_FAST_DIV
    unsigned long long *bad, hb = 0; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    div_check<<<sms * 8, 256>>>(bad);
    cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("exhaustive division check over all 2^32 fp32 inputs x 3 divisors: %llu mismatches vs IEEE division\n", hb);
#endif
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
