// Stand-alone check of csrc/tcc_mma.cuh against fp64 on the host: the three contraction shapes of the tensor-core cluster
// kernel, each as 3xTF32 (hi/lo split, two TMEM accumulators), all operands in SWIZZLE_128B shared-memory blocks.
//   test 1  forward        D[64 x 64]   = A[64 x 128] . B[64 x 128]^T      both K-major, M = 64 (16 lanes per TMEM subpartition)
//   test 2  weight grad    D[256 x 64]  = H1[64 x 256]^T . dZ2[64 x 64]    both MN-major views of forward-layout blocks, M = 128 x 2
//   test 3  input grad     D[64 x 256]  = dZ2[64 x 64] . W2T[256 x 64]^T   K-major, M = 64, N = 128 x 2
// Prints the relative L2 error of each and the SM cycles from first MMA issue to accumulator-ready.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/tcc_probe tools/tcc_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/tcc_mma.cuh"
using namespace hirl;

// src [rows][ld] fp32 -> hi / lo block sequences (cols / 32 blocks of rows x 128 B each)
__device__ void fill_blocks(char *hi, char *lo, const float *src, int ld, int rows, int cols) {
    for (int e = threadIdx.x; e < rows * cols; e += blockDim.x) {
        const int r = e / cols, c = e % cols;
        const float x = src[(size_t)r * ld + c], h = tcc::tf32_hi(x);
        const uint32_t off = (uint32_t)(c >> 5) * rows * 128 + tcc::sw128_off(r, c & 31);
        *reinterpret_cast<float *>(hi + off) = h;
        *reinterpret_cast<float *>(lo + off) = x - h;
    }
}

__global__ void __launch_bounds__(256, 1) k_probe(int test, const float *A, const float *B, float *out, long long *cycles) {
    extern __shared__ char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    char *sm = (char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = warp & 3, wh = warp >> 2;
    if (warp == 0) tc::tmem_alloc(&tmem_slot);
    if (tid == 0) { tc::mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = tmem_slot;
    char *a_hi, *a_lo, *b_hi, *b_lo;
    if (test == 1) {
        a_hi = sm; a_lo = sm + 32768; b_hi = sm + 65536; b_lo = sm + 98304;
        fill_blocks(a_hi, a_lo, A, 128, 64, 128);
        fill_blocks(b_hi, b_lo, B, 128, 64, 128);
    } else if (test == 2) {
        a_hi = sm; a_lo = sm + 65536; b_hi = sm + 131072; b_lo = sm + 147456;
        fill_blocks(a_hi, a_lo, A, 256, 64, 256);   // H1 [64][256]
        fill_blocks(b_hi, b_lo, B, 64, 64, 64);     // dZ2 [64][64]
    } else {
        a_hi = sm; a_lo = sm + 16384; b_hi = sm + 32768; b_lo = sm + 98304;
        fill_blocks(a_hi, a_lo, A, 64, 64, 64);     // dZ2 [64][64]
        fill_blocks(b_hi, b_lo, B, 64, 256, 64);    // W2T [256][64]
    }
    tc::fence_async_smem();
    __syncthreads();
    long long t0 = 0;
    if (tid == 0) {
        t0 = clock64();
        const uint32_t ah = tcc::smem_u32(a_hi), al = tcc::smem_u32(a_lo), bh = tcc::smem_u32(b_hi), bl = tcc::smem_u32(b_lo);
        if (test == 1) {
            const uint32_t id = tcc::idesc(64, 64, 0, 0);
            for (int c = 0; c < 4; ++c)
                tcc::mma3(tm + 0, tm + 64, tcc::desc_k(ah + c * 8192), tcc::desc_k(al + c * 8192), tcc::desc_k(bh + c * 8192), tcc::desc_k(bl + c * 8192), 4, 2, 2, id, c == 0);
        } else if (test == 2) {
            const uint32_t id = tcc::idesc(128, 64, 1, 1);
            for (int mb = 0; mb < 2; ++mb)
                tcc::mma3(tm + mb * 128, tm + mb * 128 + 64, tcc::desc_mn(ah + mb * 4 * 8192, 8192), tcc::desc_mn(al + mb * 4 * 8192, 8192), tcc::desc_mn(bh, 8192),
                          tcc::desc_mn(bl, 8192), 8, 64, 64, id, true);
        } else {
            const uint32_t id = tcc::idesc(64, 128, 0, 0);
            for (int nh = 0; nh < 2; ++nh)
                for (int c = 0; c < 2; ++c)
                    tcc::mma3(tm + nh * 128, tm + 256 + nh * 128, tcc::desc_k(ah + c * 8192), tcc::desc_k(al + c * 8192), tcc::desc_k(bh + c * 32768 + nh * 16384),
                              tcc::desc_k(bl + c * 32768 + nh * 16384), 4, 2, 2, id, c == 0);
        }
        tc::mma_commit(&bar);
    }
    __syncwarp();
    tc::mbar_wait(&bar, 0, 1);
    tc::tc_fence_after();
    if (tid == 0) cycles[0] = clock64() - t0;
    const uint32_t lane_addr = tm + ((uint32_t)(sub * 32) << 16);
    float m[32], c[32];
    if (test == 1) {
        tcc::tmem_ld32(lane_addr + wh * 32, m);
        tcc::tmem_ld32(lane_addr + 64 + wh * 32, c);
        if (lane < 16)
            for (int i = 0; i < 32; ++i) out[(sub * 16 + lane) * 64 + wh * 32 + i] = m[i] + c[i];
    } else if (test == 2) {
        for (int mb = 0; mb < 2; ++mb) {
            tcc::tmem_ld32(lane_addr + mb * 128 + wh * 32, m);
            tcc::tmem_ld32(lane_addr + mb * 128 + 64 + wh * 32, c);
            for (int i = 0; i < 32; ++i) out[(mb * 128 + sub * 32 + lane) * 64 + wh * 32 + i] = m[i] + c[i];
        }
    } else {
        for (int q = 0; q < 4; ++q) {
            const int col = wh * 128 + q * 32;
            tcc::tmem_ld32(lane_addr + col, m);
            tcc::tmem_ld32(lane_addr + 256 + col, c);
            if (lane < 16)
                for (int i = 0; i < 32; ++i) out[(sub * 16 + lane) * 256 + col + i] = m[i] + c[i];
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tm);
}

static double check(const char *name, const std::vector<float> &got, const std::vector<double> &ref, long long cyc) {
    double num = 0, den = 0, worst = 0;
    size_t wi = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        const double d = got[i] - ref[i];
        num += d * d; den += ref[i] * ref[i];
        if (fabs(d) > worst) { worst = fabs(d); wi = i; }
    }
    const double rel = sqrt(num / den);
    printf("%-28s rel-L2 %.3e  max|err| %.3e at %zu (got %.6f ref %.6f)  mma cycles %lld  %s\n", name, rel, worst, wi, got[wi], ref[wi], cyc,
           rel < 5e-6 ? "OK" : "MISMATCH");
    return rel;
}

int main() {
    srand(7);
    auto rnd = [](std::vector<float> &v, float s) { for (auto &x : v) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f * s; };
    float *dA, *dB, *dO;
    long long *dC, hc = 0;
    cudaMalloc(&dA, 1 << 20); cudaMalloc(&dB, 1 << 20); cudaMalloc(&dO, 1 << 20); cudaMalloc(&dC, 64);
    const int SMEM = 200 * 1024;
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    int bad = 0;
    for (int test = 1; test <= 3; ++test) {
        const int ar = 64, ac = test == 1 ? 128 : test == 2 ? 256 : 64;
        const int br = test == 3 ? 256 : 64, bc = test == 1 ? 128 : 64;
        std::vector<float> A((size_t)ar * ac), B((size_t)br * bc);
        rnd(A, 1.f); rnd(B, 0.3f);
        const int om = test == 2 ? 256 : 64, on = test == 3 ? 256 : 64;
        std::vector<double> ref((size_t)om * on, 0.0);
        if (test == 1) { for (int i = 0; i < 64; ++i) for (int j = 0; j < 64; ++j) { double s = 0; for (int k = 0; k < 128; ++k) s += (double)A[i * 128 + k] * B[j * 128 + k]; ref[i * 64 + j] = s; } }
        if (test == 2) { for (int m = 0; m < 256; ++m) for (int n = 0; n < 64; ++n) { double s = 0; for (int b = 0; b < 64; ++b) s += (double)A[b * 256 + m] * B[b * 64 + n]; ref[m * 64 + n] = s; } }
        if (test == 3) { for (int i = 0; i < 64; ++i) for (int n = 0; n < 256; ++n) { double s = 0; for (int k = 0; k < 64; ++k) s += (double)A[i * 64 + k] * B[n * 64 + k]; ref[i * 256 + n] = s; } }
        cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(dO, 0xff, 1 << 20);
        for (int rep = 0; rep < 2; ++rep) k_probe<<<1, 256, SMEM>>>(test, dA, dB, dO, dC);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("test %d: %s\n", test, cudaGetErrorString(e)); return 1; }
        std::vector<float> got(ref.size());
        cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&hc, dC, 8, cudaMemcpyDeviceToHost);
        const char *names[] = {"", "fwd K-major M=64 N=64", "dW2 MN-major M=128x2 N=64", "dH1 K-major M=64 N=128x2"};
        if (check(names[test], got, ref, hc) >= 5e-6) ++bad;
    }
    printf(bad ? "tcc_probe: %d test(s) FAILED\n" : "tcc_probe: all ok\n", bad);
    return bad ? 2 : 0;
}
