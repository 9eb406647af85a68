// Throughput probe for design decisions (not part of the product).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) ffma_imm(float *out, int iters, float a, float b) {
    float x[16];
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 1.2345f) out[0] = s;
}
// GEMM-like: 4x4 outer product per step, a[4], b[4] change every step (register operands)
__global__ void __launch_bounds__(256) ffma_outer(float *out, int iters, const float *src) {
    float acc[16], a[4], b[4];
    for (int i = 0; i < 16; ++i) acc[i] = 0;
    for (int i = 0; i < 4; ++i) { a[i] = src[threadIdx.x + i]; b[i] = src[threadIdx.x + 4 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i * 4 + j] = fmaf(a[i], b[j], acc[i * 4 + j]);
        // rotate operands cheaply so the compiler cannot hoist
        float t = a[0]; a[0] = a[1]; a[1] = a[2]; a[2] = a[3]; a[3] = b[0]; b[0] = b[1]; b[1] = b[2]; b[2] = b[3]; b[3] = t;
    }
    float s = 0; for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 1.2345f) out[0] = s;
}
// 8x8 register tile
__global__ void __launch_bounds__(256) ffma_outer8(float *out, int iters, const float *src) {
    float acc[64], a[8], b[8];
    for (int i = 0; i < 64; ++i) acc[i] = 0;
    for (int i = 0; i < 8; ++i) { a[i] = src[threadIdx.x + i]; b[i] = src[threadIdx.x + 8 + i]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(a[i], b[j], acc[i * 8 + j]);
        float t = a[0];
#pragma unroll
        for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
        a[7] = b[0];
#pragma unroll
        for (int i = 0; i < 7; ++i) b[i] = b[i + 1];
        b[7] = t;
    }
    float s = 0; for (int i = 0; i < 64; ++i) s += acc[i];
    if (s == 1.2345f) out[0] = s;
}
// legacy tensor path: mma.sync m16n8k8 tf32, 8 independent accumulators per warp
__global__ void __launch_bounds__(256) mma_tf32(float *out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1.2345f) out[0] = s;
}
__global__ void __launch_bounds__(256) mma_bf16(float *out, int iters) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 + 4, b1 = a0 + 5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
    float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 1.2345f) out[0] = s;
}
template <class F> double timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms * 1e-3;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float *out, *src; cudaMalloc(&out, 4096); cudaMalloc(&src, 4096); cudaMemset(src, 0, 4096);
    const int iters = 1 << 14;
    for (int wpb : {8}) for (int bps : {1, 2, 4, 8}) {
        int blocks = sms * bps;
        double t;
        t = timeit([&] { ffma_imm<<<blocks, 256>>>(out, iters, 0.999f, 0.001f); });
        printf("blocks/SM %d: ffma_imm   %7.2f TFLOP/s\n", bps, 2.0 * 16 * iters * 256.0 * blocks / t / 1e12);
        t = timeit([&] { ffma_outer<<<blocks, 256>>>(out, iters, src); });
        printf("blocks/SM %d: ffma_outer4 %7.2f TFLOP/s\n", bps, 2.0 * 16 * iters * 256.0 * blocks / t / 1e12);
        t = timeit([&] { ffma_outer8<<<blocks, 256>>>(out, iters / 4, src); });
        printf("blocks/SM %d: ffma_outer8 %7.2f TFLOP/s\n", bps, 2.0 * 64 * (iters / 4) * 256.0 * blocks / t / 1e12);
        t = timeit([&] { mma_tf32<<<blocks, 256>>>(out, iters / 4); });
        printf("blocks/SM %d: mma.sync tf32 m16n8k8  %7.2f TFLOP/s\n", bps, 2.0 * 16 * 8 * 8 * 8 * (iters / 4) * 8.0 * blocks / t / 1e12);
        t = timeit([&] { mma_bf16<<<blocks, 256>>>(out, iters / 4); });
        printf("blocks/SM %d: mma.sync bf16 m16n8k16 %7.2f TFLOP/s\n", bps, 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * 8.0 * blocks / t / 1e12);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
