// Standalone timing probe of the FP32 64x64 tile function of csrc/update.cuh (and experimental variants):
// every CTA runs one tile `reps` times on L2-resident operands, in-kernel clock64 per call.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/fp32_gemm_probe tools/fp32_gemm_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../hirl4ucav_b200/csrc/update.cuh"
using namespace hirl;

// ---- variant 1: conflict-free transposing store (a warp covers 32 consecutive rows of one k-quad) ----
__device__ __forceinline__ float4 load_kmajor_v1(const float *P, int ld, int r0, int R, int k0, int K, int tid) {
    const int r = r0 + (tid & 63), k = k0 + 4 * (tid >> 6);
    if (r < R && k < K) return ldtile4(P + (size_t)r * ld + k);
    return make_float4(0.f, 0.f, 0.f, 0.f);
}
__device__ __forceinline__ void store_kmajor_v1(float *T, float4 v, int tid) {
    const int r = tid & 63, kq = 4 * (tid >> 6);
    T[(kq + 0) * TS + r] = v.x;
    T[(kq + 1) * TS + r] = v.y;
    T[(kq + 2) * TS + r] = v.z;
    T[(kq + 3) * TS + r] = v.w;
}

template <int VAR>
__device__ __noinline__ void tile_nt(const float *Ap, int lda, const float *Bp, int ldb, float *C, int ldc, int M, int N, int K, int m0,
                                     int n0, int kt_begin, int kt_end, float *smem) {
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float *As = smem, *Bs = smem + 2 * 16 * TS;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#define FETCH(kt, ra, rb)                                                                      \
    {                                                                                          \
        const int k0_ = (kt) << 4;                                                             \
        if (VAR == 1) { ra = load_kmajor_v1(Ap, lda, m0, M, k0_, K, tid); rb = load_kmajor_v1(Bp, ldb, n0, N, k0_, K, tid); } \
        else { ra = load_kmajor(Ap, lda, m0, M, k0_, K, tid); rb = load_kmajor(Bp, ldb, n0, N, k0_, K, tid); }               \
    }
#define STASH(buf, ra, rb)                                                                     \
    {                                                                                          \
        if (VAR == 2) { /* ablation: no shared-memory stores */ }                              \
        else if (VAR == 1) { store_kmajor_v1(As + (buf) * 16 * TS, ra, tid); store_kmajor_v1(Bs + (buf) * 16 * TS, rb, tid); } \
        else { store_kmajor(As + (buf) * 16 * TS, ra, tid); store_kmajor(Bs + (buf) * 16 * TS, rb, tid); }                    \
    }
#define COMPUTE(buf)                                                                             \
    {                                                                                            \
        const float *a_ = As + (buf) * 16 * TS + ty * 4, *b_ = Bs + (buf) * 16 * TS + tx * 4;    \
        float4 av = ld4(a_), bv = ld4(b_);                                                       \
        _Pragma("unroll") for (int kk = 0; kk < 16; ++kk) {                                      \
            float4 an = av, bn = bv;                                                             \
            if (kk + 1 < 16) { an = ld4(a_ + (kk + 1) * TS); bn = ld4(b_ + (kk + 1) * TS); }     \
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};      \
            _Pragma("unroll") for (int i = 0; i < 4; ++i)                                        \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]); \
            av = an; bv = bn;                                                                    \
        }                                                                                        \
    }
    float4 ra0, rb0, ra1, rb1;
    __syncthreads();
    if (kt_begin < kt_end) {
        FETCH(kt_begin, ra0, rb0)
        STASH(0, ra0, rb0)
        if (kt_begin + 1 < kt_end) FETCH(kt_begin + 1, ra1, rb1)
    }
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; kt += 2) {
        if (VAR != 3 && kt + 2 < kt_end) FETCH(kt + 2, ra0, rb0)
        COMPUTE(0)
        if (kt + 1 < kt_end) STASH(1, ra1, rb1)
        __syncthreads();
        if (kt + 1 >= kt_end) break;
        if (VAR != 3 && kt + 3 < kt_end) FETCH(kt + 3, ra1, rb1)
        COMPUTE(1)
        if (kt + 2 < kt_end) STASH(0, ra0, rb0)
        __syncthreads();
    }
    const int n = n0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m < M && n < N) st4(C + (size_t)m * ldc + n, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
}


// ---- variant 4: cp.async ring (3 stages), operands kept k-contiguous in shared memory ([row][16] with a 16-byte
// chunk swizzle), no register staging and no transposing stores; thread (tx,ty) owns rows ty+16i, cols tx+16j ----
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int ABL>
__device__ __noinline__ void tile_nt_cpasync(const float *Ap, int lda, const float *Bp, int ldb, float *C, int ldc, int M, int N, int K,
                                             int m0, int n0, int kt_begin, int kt_end, float *smem) {
    constexpr int STG = 3, OPF = 64 * 16;   // floats per operand per stage
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // copy role: row r = tid>>2, 16-byte chunk c = tid&3
    const int cr = tid >> 2, cc = tid & 3;
    const int cdst = cr * 16 + 4 * (cc ^ ((cr >> 1) & 3));
    const float *asrc = Ap + (size_t)(m0 + cr) * lda + 4 * cc, *bsrc = Bp + (size_t)(n0 + cr) * ldb + 4 * cc;
    const bool aval = m0 + cr < M, bval = n0 + cr < N;
    auto issue = [&](int kt) {
        const int st = (kt - kt_begin) % STG;
        cp_async16(smem + (2 * st) * OPF + cdst, asrc + (kt << 4), aval);
        cp_async16(smem + (2 * st + 1) * OPF + cdst, bsrc + (kt << 4), bval);
    };
    __syncthreads();
    if (kt_begin < kt_end) issue(kt_begin);
    cp_commit();
    if (kt_begin + 1 < kt_end) issue(kt_begin + 1);
    cp_commit();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
        cp_wait<1>();
        __syncthreads();
        if (ABL != 1 && kt + 2 < kt_end) issue(kt + 2);
        cp_commit();
        const int st = (kt - kt_begin) % STG;
        const float *As = smem + (2 * st) * OPF, *Bs = smem + (2 * st + 1) * OPF;
#pragma unroll
        for (int c = 0; c < 4; ++c) {   // 4 k-steps per 16-byte chunk
            float4 a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = ty + 16 * i;
                a[i] = ABL == 2 ? make_float4(1.f + c, 2.f, 3.f, 4.f + i) : ld4(As + r * 16 + 4 * (c ^ ((r >> 1) & 3)));
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = tx + 16 * j;
                b[j] = ABL == 2 ? make_float4(1.f, 2.f + c, 3.f + j, 4.f) : ld4(Bs + r * 16 + 4 * (c ^ ((r >> 1) & 3)));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j];
        }
    }
}

// ---- variant 7: as variant 4, fragments double-buffered explicitly (chunk c+1 is loaded before chunk c is used) ----
__device__ __noinline__ void tile_nt_cpasync_db(const float *Ap, int lda, const float *Bp, int ldb, float *C, int ldc, int M, int N, int K,
                                                int m0, int n0, int kt_begin, int kt_end, float *smem) {
    constexpr int STG = 3, OPF = 64 * 16;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int cr = tid >> 2, cc = tid & 3;
    const int cdst = cr * 16 + 4 * (cc ^ ((cr >> 1) & 3));
    const float *asrc = Ap + (size_t)(m0 + cr) * lda + 4 * cc, *bsrc = Bp + (size_t)(n0 + cr) * ldb + 4 * cc;
    const bool aval = m0 + cr < M, bval = n0 + cr < N;
    auto issue = [&](int kt) {
        const int st = (kt - kt_begin) % STG;
        cp_async16(smem + (2 * st) * OPF + cdst, asrc + (kt << 4), aval);
        cp_async16(smem + (2 * st + 1) * OPF + cdst, bsrc + (kt << 4), bval);
    };
    // per-thread fragment offsets (floats) for chunk 0; chunk c flips the swizzled chunk index
    int aoff[4], boff[4], asw[4], bsw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * i, q = tx + 16 * i;
        aoff[i] = r * 16; asw[i] = (r >> 1) & 3;
        boff[i] = q * 16; bsw[i] = (q >> 1) & 3;
    }
    __syncthreads();
    if (kt_begin < kt_end) issue(kt_begin);
    cp_commit();
    if (kt_begin + 1 < kt_end) issue(kt_begin + 1);
    cp_commit();
    for (int kt = kt_begin; kt < kt_end; ++kt) {
        cp_wait<1>();
        __syncthreads();
        if (kt + 2 < kt_end) issue(kt + 2);
        cp_commit();
        const int st = (kt - kt_begin) % STG;
        const float *As = smem + (2 * st) * OPF, *Bs = smem + (2 * st + 1) * OPF;
        float4 a[2][4], b[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[0][i] = ld4(As + aoff[i] + 4 * (0 ^ asw[i])); b[0][i] = ld4(Bs + boff[i] + 4 * (0 ^ bsw[i])); }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int cur = c & 1, nxt = cur ^ 1;
            if (c + 1 < 4) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a[nxt][i] = ld4(As + aoff[i] + 4 * ((c + 1) ^ asw[i]));
                    b[nxt][i] = ld4(Bs + boff[i] + 4 * ((c + 1) ^ bsw[i]));
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(a[cur][i].x, b[cur][j].x, acc[i][j]);
                    acc[i][j] = fmaf(a[cur][i].y, b[cur][j].y, acc[i][j]);
                    acc[i][j] = fmaf(a[cur][i].z, b[cur][j].z, acc[i][j]);
                    acc[i][j] = fmaf(a[cur][i].w, b[cur][j].w, acc[i][j]);
                }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j];
        }
    }
}

// ---- variant 8: round-1 pipeline (register prefetch two k-tiles ahead, double-buffered smem) with the k-contiguous
// swizzled layout: one LDG.128 + one STS.128 per operand, thread and k-tile; fragment offsets precomputed ----
__device__ __noinline__ void tile_nt_regs_kc(const float *Ap, int lda, const float *Bp, int ldb, float *C, int ldc, const float *bias,
                                             int M, int N, int K, int m0, int n0, int kt_begin, int kt_end, float *smem) {
    constexpr int OPF = 64 * 16;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int cr = tid >> 2, cc = tid & 3;
    const int cdst = cr * 16 + 4 * (cc ^ ((cr >> 1) & 3));
    const bool aval = m0 + cr < M, bval = n0 + cr < N;
    const float *asrc = Ap + (size_t)(aval ? m0 + cr : 0) * lda + 4 * cc, *bsrc = Bp + (size_t)(bval ? n0 + cr : 0) * ldb + 4 * cc;
    int aoff[4][4], boff[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = ty + 16 * i, q = tx + 16 * i;
#pragma unroll
        for (int c = 0; c < 4; ++c) { aoff[i][c] = r * 16 + 4 * (c ^ ((r >> 1) & 3)); boff[i][c] = q * 16 + 4 * (c ^ ((q >> 1) & 3)); }
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#define FETCH8(kt, ra, rb) { ra = aval ? ldtile4(asrc + ((kt) << 4)) : zero4; rb = bval ? ldtile4(bsrc + ((kt) << 4)) : zero4; }
#define STASH8(buf, ra, rb) { st4(smem + (2 * (buf)) * OPF + cdst, ra); st4(smem + (2 * (buf) + 1) * OPF + cdst, rb); }
#define COMPUTE8(buf)                                                                                      \
    {                                                                                                      \
        const float *As = smem + (2 * (buf)) * OPF, *Bs = As + OPF;                                        \
        _Pragma("unroll") for (int c = 0; c < 4; ++c) {                                                    \
            float4 a[4], b[4];                                                                             \
            _Pragma("unroll") for (int i = 0; i < 4; ++i) { a[i] = ld4(As + aoff[i][c]); b[i] = ld4(Bs + boff[i][c]); } \
            _Pragma("unroll") for (int i = 0; i < 4; ++i)                                                  \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) {                                            \
                    acc[i][j] = fmaf(a[i].x, b[j].x, acc[i][j]);                                           \
                    acc[i][j] = fmaf(a[i].y, b[j].y, acc[i][j]);                                           \
                    acc[i][j] = fmaf(a[i].z, b[j].z, acc[i][j]);                                           \
                    acc[i][j] = fmaf(a[i].w, b[j].w, acc[i][j]);                                           \
                }                                                                                          \
        }                                                                                                  \
    }
    float4 ra0, rb0, ra1, rb1;
    __syncthreads();
    if (kt_begin < kt_end) {
        FETCH8(kt_begin, ra0, rb0)
        STASH8(0, ra0, rb0)
        if (kt_begin + 1 < kt_end) FETCH8(kt_begin + 1, ra1, rb1)
    }
    __syncthreads();
    for (int kt = kt_begin; kt < kt_end; kt += 2) {
        if (kt + 2 < kt_end) FETCH8(kt + 2, ra0, rb0)
        COMPUTE8(0)
        if (kt + 1 < kt_end) STASH8(1, ra1, rb1)
        __syncthreads();
        if (kt + 1 >= kt_end) break;
        if (kt + 3 < kt_end) FETCH8(kt + 3, ra1, rb1)
        COMPUTE8(1)
        if (kt + 2 < kt_end) STASH8(0, ra0, rb0)
        __syncthreads();
    }
    float bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bj[j] = (bias && n0 + tx + 16 * j < N) ? bias[n0 + tx + 16 * j] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty + 16 * i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx + 16 * j;
            if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j] + bj[j];
        }
    }
}

template <int VAR>
__global__ void __launch_bounds__(256, 1) k_probe(const float *A, const float *B, float *C, int M, int N, int K, int reps, long long *cyc, int nkt, const float *bias) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    const int ntn = N / 64, tiles = (M / 64) * ntn;
    const int t = blockIdx.x % tiles;
    long long t0 = clock64(), tfirst = 0;
    for (int r = 0; r < reps; ++r) {
        if (r == 1) tfirst = clock64() - t0;
        if (VAR == 8) tile_nt_regs_kc(A, K, B, K, C, N, bias, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else if (VAR == 7) tile_nt_cpasync_db(A, K, B, K, C, N, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else if (VAR == 5) tile_nt_cpasync<1>(A, K, B, K, C, N, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else if (VAR == 6) tile_nt_cpasync<2>(A, K, B, K, C, N, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else if (VAR == 4) tile_nt_cpasync<0>(A, K, B, K, C, N, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else if (VAR == 9) gemm_tile<GEMM_NT>(A, K, B, K, C, N, bias, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
        else tile_nt<VAR>(A, K, B, K, C, N, M, N, K, (t / ntn) * 64, (t % ntn) * 64, 0, nkt, smem);
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[blockIdx.x] = (t1 - t0 - tfirst) / (reps > 1 ? reps - 1 : 1); cyc[148 + blockIdx.x] = tfirst; }
}

static int g_nkt = 0;
static const float *g_bias = nullptr;
template <int VAR>
static void run(const char *name, const float *dA, const float *dB, float *dC, int M, int N, int K, long long *dcyc) {
    cudaFuncSetAttribute(k_probe<VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    k_probe<VAR><<<148, 256, SMEM_BYTES>>>(dA, dB, dC, M, N, K, 50, dcyc, g_nkt ? g_nkt : K / 16, g_bias);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(296);
    cudaMemcpy(h.data(), dcyc, 296 * 8, cudaMemcpyDeviceToHost);
    double s = 0, mx = 0, sf = 0;
    for (int i = 0; i < 148; ++i) { s += h[i]; if (h[i] > mx) mx = h[i]; sf += h[148 + i]; }
    std::vector<float> hc(16);
    cudaMemcpy(hc.data(), dC, 64, cudaMemcpyDeviceToHost);
    printf("%-34s %s  mean %.0f cyc/tile (%.2f us @1.965GHz), max %.0f, per k-tile %.0f cyc, FIRST (cold) call %.0f cyc  C[0]=%g\n", name, cudaGetErrorString(e), s / 148,
           s / 148 / 1965.0, mx, s / 148 / (g_nkt ? g_nkt : K / 16), sf / 148, hc[0]);
}

int main(int argc, char **argv) {
    int M = 256, N = 512, K = argc > 1 ? atoi(argv[1]) : 256;
    std::vector<float> hA((size_t)M * K), hB((size_t)N * K);
    srand(1);
    for (auto &x : hA) x = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto &x : hB) x = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    float *dA, *dB, *dC;
    long long *dcyc;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dC, (size_t)M * N * 4); cudaMalloc(&dcyc, 296 * 8);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    g_nkt = argc > 2 ? atoi(argv[2]) : 0;
    if (argc > 3) { float *b; cudaMalloc(&b, N * 4); cudaMemset(b, 0, N * 4); g_bias = b; }
    run<9>("gemm_tile<NT> (cp.async ring)", dA, dB, dC, M, N, K, dcyc);
    run<8>("register staged, k-contiguous smem", dA, dB, dC, M, N, K, dcyc);
    run<0>("gemm_tile_regs<NT> (product, copy)", dA, dB, dC, M, N, K, dcyc);
    run<1>("conflict-free transposing STS", dA, dB, dC, M, N, K, dcyc);
    run<2>("ablation: no STS (wrong result)", dA, dB, dC, M, N, K, dcyc);
    run<3>("ablation: no LDG in the loop", dA, dB, dC, M, N, K, dcyc);
    run<4>("cp.async ring, k-contiguous smem", dA, dB, dC, M, N, K, dcyc);
    run<7>("cp.async ring, fragments double-buffered", dA, dB, dC, M, N, K, dcyc);
    run<5>("  ablation: no cp.async in loop", dA, dB, dC, M, N, K, dcyc);
    run<6>("  ablation: no LDS (const operands)", dA, dB, dC, M, N, K, dcyc);
    return 0;
}
