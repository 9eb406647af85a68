// env.cuh -- batched Harfang observation / reward / lock-launch flag math
// (reference hirl/environments/HarfangEnv_GYM.py:101-137, 160-169, 193-268, 299-336).
//
// Bandwidth-bound elementwise pass: 64 B in, 60 B out per state.  The reference does this
// arithmetic on Python floats, so every operation here is IEEE fp64 on the widened fp32
// inputs, in the reference's order, and only the final observation/reward are rounded to
// fp32 (what the replay buffer stores, HIRL.py:229-233).  Thresholds are compared in fp64:
// for fp32 x,  (double)x <= 0.1  is not  x <= 0.1f.
#pragma once
#include "common.cuh"

namespace hirl {

constexpr int ENV_THREADS = 256;
#ifndef HIRL_ENV_MIN_BLOCKS
#define HIRL_ENV_MIN_BLOCKS 6
#endif
#ifndef HIRL_ENV_FAST_DIV
#define HIRL_ENV_FAST_DIV 1
#endif

// x / d for a compile-time-known divisor with reciprocal r = RN(1/d): q0 = x*r, one FMA residual correction
// (Markstein), and the sign of zero taken from q0.  The generic IEEE fp64 division sequence (11 of them per
// state) made the kernel FP64-issue bound at 73 % of HBM peak; this form reaches 87 %.  It is NOT an
// approximation here: tools/env_bench.cu compares it with IEEE division for ALL 2^32 fp32 bit patterns x
// (the only inputs this kernel ever sees) and the three divisors used below -- 0 mismatches.
__device__ __forceinline__ double div_const(double x, double d, double r) {
#if HIRL_ENV_FAST_DIV
    const double q0 = x * r;
    const double rem = fma(-q0, d, x);
    const double q = fma(rem, r, q0);
    return x == 0.0 ? q0 : q;      // keeps -0.0 / d == -0.0
#else
    return x / d;
#endif
}

struct EnvOut { float obs[13]; float reward; int flags; };

__device__ __forceinline__ void env_one(const float4 &q0, const float4 &q1, const float4 &q2, const float4 &q3, EnvOut &o) {
    const double pax = q0.x, pay = q0.y, paz = q0.z, eax = q0.w, eay = q1.x, eaz = q1.y;
    const double pox = q1.z, poy = q1.w, poz = q2.x, eox = q2.y, eoy = q2.z, eoz = q2.w;
    const double angle = q3.x, health = q3.y, fire_cmd = q3.z;
    const unsigned fl = __float_as_uint(q3.w);
    const bool locked_prev = fl & 1u, locked_now = fl & 2u, missile_prev = fl & 4u, missile_now = fl & 8u;
    bool fire_success = fl & 16u;
    const double PI_D = 3.141592653589793;
    // observation (GYM.py:196-201, 213-218, 232, 262): normalise both, then subtract
    const double R_POS = 1.0 / 10000.0, R_PI = 1.0 / 3.141592653589793, R_180 = 1.0 / 180.0;
    o.obs[0] = (float)(div_const(pax, 10000.0, R_POS) - div_const(pox, 10000.0, R_POS));
    o.obs[1] = (float)(div_const(pay, 10000.0, R_POS) - div_const(poy, 10000.0, R_POS));
    o.obs[2] = (float)(div_const(paz, 10000.0, R_POS) - div_const(poz, 10000.0, R_POS));
    o.obs[3] = (float)div_const(eax, PI_D, R_PI);
    o.obs[4] = (float)div_const(eay, PI_D, R_PI);
    o.obs[5] = (float)div_const(eaz, PI_D, R_PI);
    const double t_angle = div_const(angle, 180.0, R_180);
    o.obs[6] = (float)t_angle;
    o.obs[7] = locked_now ? 1.f : -1.f;
    o.obs[8] = missile_now ? 1.f : -1.f;
    o.obs[9] = (float)div_const(eox, PI_D, R_PI);
    o.obs[10] = (float)div_const(eoy, PI_D, R_PI);
    o.obs[11] = (float)div_const(eoz, PI_D, R_PI);
    o.obs[12] = (float)health;
    // reward (GYM.py:101-137), sequential -= like the reference
    const double dx = pax - pox, dy = pay - poy, dz = paz - poz;
    const double loc = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
    const bool fired = fire_cmd > 0.0;  // GYM.py:150
    double r = 0.0;
    r = __dsub_rn(r, __dmul_rn(0.0001, loc));
    r = __dsub_rn(r, __dmul_rn(t_angle, 10.0));
    if (pay < 2000.0) r -= 4.0;
    if (pay > 7000.0) r -= 4.0;
    int succ = 0;
    if (fired) {
        r -= 8.0;
        if (missile_prev && !locked_prev) succ = -1;
        else if (missile_prev && locked_prev) { succ = 1; fire_success = true; }
    }
    if (health <= 0.1 && fire_success) r += 600.0;
    o.reward = (float)r;
    // termination (GYM.py:160-169)
    bool done = (pay < 500.0) || (pay > 10000.0);
    const bool ep = health <= 0.0;
    done = done || ep;
    // bit 3: the policy's fire rule a[3] > 0 (GYM.py:150); bit 4: the EXPERT's fire rule on the new observation,
    // obs[7] > 0 and obs[8] > 0 = target locked and missile slot loaded (data/straight_line/ai_data_col.py:63-64)
    const bool expert_fire = locked_now && missile_now;
    o.flags = (int)done | ((int)ep << 1) | ((int)fire_success << 2) | ((int)fired << 3) | ((int)expert_fire << 4) | ((succ & 0xFF) << 8);
}

// one thread per state; outputs are staged through shared memory so that every global store is
// a full 128-bit coalesced transaction (13-float rows are not 16 B aligned per state).
__global__ void __launch_bounds__(ENV_THREADS, HIRL_ENV_MIN_BLOCKS) env_step_kernel(const float4 *__restrict__ raw, long long n, float *__restrict__ obs,
                                                               float *__restrict__ reward, int *__restrict__ flags,
                                                               unsigned long long *__restrict__ counters) {
    __shared__ __align__(16) float s_obs[ENV_THREADS * 13];
    __shared__ unsigned s_cnt[3];
    const int tid = threadIdx.x;
    if (counters && tid < 3) s_cnt[tid] = 0u;
    unsigned c_fire = 0, c_succ = 0, c_done = 0;
    const long long tiles = (n + ENV_THREADS - 1) / ENV_THREADS;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const long long base = tile * ENV_THREADS, i = base + tid;
        EnvOut o;
        const bool valid = i < n;
        if (valid) {
            const float4 q0 = __ldcs(raw + 4 * i), q1 = __ldcs(raw + 4 * i + 1), q2 = __ldcs(raw + 4 * i + 2), q3 = __ldcs(raw + 4 * i + 3);
            env_one(q0, q1, q2, q3, o);
            __stcs(reward + i, o.reward);
            __stcs(flags + i, o.flags);
            const int succ = (int)(signed char)((o.flags >> 8) & 0xFF);
            c_fire += succ != 0;
            c_succ += succ == 1;
            c_done += o.flags & 1;
#pragma unroll
            for (int k = 0; k < 13; ++k) s_obs[tid * 13 + k] = o.obs[k];
        }
        __syncthreads();
        const long long cnt = (n - base < ENV_THREADS ? n - base : ENV_THREADS) * 13;  // floats in this tile
        float *dst = obs + base * 13;  // base*13*4 bytes: 256*13*4 = 13312 B per tile -> 16 B aligned
        const long long vec = cnt >> 2;
        for (long long v = tid; v < vec; v += ENV_THREADS) __stcs(reinterpret_cast<float4 *>(dst) + v, reinterpret_cast<const float4 *>(s_obs)[v]);
        for (long long v = (vec << 2) + tid; v < cnt; v += ENV_THREADS) dst[v] = s_obs[v];
        __syncthreads();
    }
    if (counters) {
        c_fire = __reduce_add_sync(0xffffffffu, c_fire);
        c_succ = __reduce_add_sync(0xffffffffu, c_succ);
        c_done = __reduce_add_sync(0xffffffffu, c_done);
        if ((tid & 31) == 0) { atomicAdd(&s_cnt[0], c_fire); atomicAdd(&s_cnt[1], c_succ); atomicAdd(&s_cnt[2], c_done); }
        __syncthreads();
        if (tid < 3 && s_cnt[tid]) atomicAdd(counters + tid, (unsigned long long)s_cnt[tid]);
    }
}

// offline reward / termination on normalised observations (GYM.py:299-336).  T = float: replay-shaped fp32 rows;
// T = double: the expert CSV as read_data delivers it (utils/data_processor.py:5-11) -- the reference evaluates these
// thresholds on the float64 values, and (double)(float)x <= 0.1 is not x <= 0.1 for x = 0.1.
template <typename T>
__global__ void __launch_bounds__(ENV_THREADS) env_offline_kernel(const T *__restrict__ s, const T *__restrict__ a,
                                                                  const T *__restrict__ s2, long long n, float *__restrict__ reward,
                                                                  signed char *__restrict__ success, unsigned char *__restrict__ done) {
    for (long long i = (long long)blockIdx.x * ENV_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * ENV_THREADS) {
        const T *p = s + 13 * i, *q = s2 + 13 * i;
        const double x = (double)q[0] * 10000.0, y = (double)q[1] * 10000.0, z = (double)q[2] * 10000.0;
        const double loc = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
        double r = 0.0;
        r = __dsub_rn(r, __dmul_rn(0.0001, loc));
        r = __dsub_rn(r, __dmul_rn((double)q[6], 10.0));
        int succ = 0;
        if ((double)a[4 * i + 3] > 0.0) {
            r -= 8.0;
            const double have = p[8], lock = p[7];
            if (have > 0.0 && lock < 0.0) succ = -1;
            else if (have > 0.0 && lock > 0.0) succ = 1;
        }
        const double hp = q[12];
        if (hp < 0.1) r += 600.0;
        reward[i] = (float)r;
        success[i] = (signed char)succ;
        done[i] = hp <= 0.1;
    }
}

}  // namespace hirl
