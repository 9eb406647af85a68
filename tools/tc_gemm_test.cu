// Standalone correctness/timing check of csrc/tc_gemm.cuh (3xTF32 tcgen05 tile) against fp64 on the host.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../hirl4ucav_b200/csrc/tc_gemm.cuh"
using namespace hirl;

__global__ void __launch_bounds__(256, 1) k_test(const float *A, int lda, int M, const float *B, int ldb, int N, int K, float *C, int ldc,
                                                  const float *bias, int reps) {
    extern __shared__ char smem_raw[];
    __shared__ tc::Barriers bars;
    __shared__ uint32_t tmem_slot;
    char *smem = (char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    if (threadIdx.x < 32) tc::tmem_alloc(&tmem_slot);
    if (threadIdx.x == 0) tc::barriers_init(&bars);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int ntn = N / tc::BN, ntm = (M + tc::BM - 1) / tc::BM;
    for (int rep = 0; rep < reps; ++rep) {
        for (int t = blockIdx.x; t < ntm * ntn; t += gridDim.x)
            tc::gemm_tile(A, lda, M, B, ldb, N, K, (t / ntn) * tc::BM, (t % ntn) * tc::BN, 0, K, C, ldc, bias, tmem_base, smem, &bars);
#ifdef TEST_FENCE_BETWEEN
        // what the persistent kernel does between two GEMM stages: row-stage smem traffic + a gpu-scope fence
        __syncthreads();
        for (int i = threadIdx.x; i < 16000; i += blockDim.x) reinterpret_cast<float *>(smem_raw)[i] = (float)i;
        __syncthreads();
        if (threadIdx.x < 32) __threadfence();
        __syncthreads();
#endif
    }
    __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tmem_base);
}

int main(int argc, char **argv) {
    int M = argc > 1 ? atoi(argv[1]) : 256, N = argc > 2 ? atoi(argv[2]) : 512, K = argc > 3 ? atoi(argv[3]) : 256;
    std::vector<float> hA((size_t)M * K), hB((size_t)N * K), hbias(N), hC((size_t)M * N);
    srand(1);
    for (auto &x : hA) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto &x : hB) x = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    for (auto &x : hbias) x = rand() / (float)RAND_MAX;
    float *dA, *dB, *dC, *dbias;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dC, hC.size() * 4); cudaMalloc(&dbias, N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, hC.size() * 4);
    cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    int tiles = ((M + 127) / 128) * (N / 64);
    int grid = tiles < 148 ? tiles : 148;
    k_test<<<grid, 256, tc::SMEM_BYTES>>>(dA, K, M, dB, K, N, K, dC, N, dbias, 1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double num = 0, den = 0, maxabs = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double s = hbias[n];
            for (int k = 0; k < K; ++k) s += (double)hA[(size_t)m * K + k] * hB[(size_t)n * K + k];
            double d = hC[(size_t)m * N + n] - s;
            num += d * d; den += s * s; if (fabs(d) > maxabs) maxabs = fabs(d);
        }
    printf("M=%d N=%d K=%d  rel-L2 err vs fp64 = %.3e  max abs = %.3e  (C[0]=%f C[last]=%f)\n", M, N, K, sqrt(num / den), maxabs, hC[0], hC.back());
    // timing
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    cudaEventRecord(e0);
    k_test<<<grid, 256, tc::SMEM_BYTES>>>(dA, K, M, dB, K, N, K, dC, N, dbias, reps);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
#if HIRL_TC_DEBUG == 4
    long long dbg[64];
    cudaMemcpyFromSymbol(dbg, tc::g_tc_dbg, sizeof(dbg));
    double ntile = (double)(reps + 1) * ((tiles + grid - 1) / grid);
    const char *nm[] = {"wait prefetched loads", "issue loads + wait empty", "split + STTM + ST B", "wait::st + fences + arrive", "issue MMAs (warp 0)", "wait chunk", "readout + sync"};
    for (int ph = 0; ph < 7; ++ph) {
        printf("  %-28s", nm[ph]);
        for (int w = 0; w < 8; ++w) printf(" %7.0f", dbg[ph * 8 + w] / ntile);
        printf("  cycles/tile per warp\n");
    }
#endif
    double per = ms * 1e3 / reps / ((tiles + grid - 1) / grid);
    printf("grid %d, %d tiles: %.2f us per tile-wave, %.1f TFLOP/s (fp32-equivalent)\n", grid, tiles, per, 2.0 * M * N * K * reps / (ms * 1e-3) / 1e12);
    return 0;
}
