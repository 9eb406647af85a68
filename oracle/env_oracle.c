/* TEST INFRASTRUCTURE -- plain-C restatement of the Harfang env arithmetic (float64).
 *
 * Same contract as oracle/env_oracle.py (which documents the raw record layout); this
 * version exists because the reference evaluates  x ** (1/2)  on Python floats, i.e.
 * libm pow(x, 0.5), which is NOT always equal to sqrt(x) (0.08 % of inputs differ by one
 * f64 ulp with glibc 2.39).  Calling the same libm pow makes the oracle bit-identical to
 * the reference at 10 M states/s instead of 0.1 M.  Also used (multi-threaded with
 * OpenMP) as the CPU baseline of the env kernel in bench.py.
 *
 * Follows (reference checkout):
 *   hirl/environments/HarfangEnv_GYM.py:193-268  observation
 *   hirl/environments/HarfangEnv_GYM.py:101-137  reward / step_success / fire_success
 *   hirl/environments/HarfangEnv_GYM.py:160-169  termination
 *   hirl/environments/HarfangEnv_GYM.py:299-336  offline reward / termination
 *   hirl/environments/constants.py:13-16         normalisers
 * Never linked into the product library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define POS_NORM 10000.0
static const double EULER_NORM = 3.141592653589793; /* math.pi */

void env_oracle_online(const float *raw, long n, double *obs, double *reward, int32_t *out_flags)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const float *q = raw + 16 * i;
        double x[15];
        for (int k = 0; k < 15; ++k) x[k] = (double)q[k];
        uint32_t fl;
        memcpy(&fl, q + 15, 4);
        int locked_prev = fl & 1, locked_now = (fl >> 1) & 1, missile_prev = (fl >> 2) & 1,
            missile_now = (fl >> 3) & 1, fire_success = (fl >> 4) & 1;
        double *o = obs + 13 * i;
        for (int k = 0; k < 3; ++k) o[k] = x[k] / POS_NORM - x[6 + k] / POS_NORM;
        for (int k = 0; k < 3; ++k) o[3 + k] = x[3 + k] / EULER_NORM;
        double t_angle = x[12] / 180;
        o[6] = t_angle;
        o[7] = locked_now ? 1.0 : -1.0;
        o[8] = missile_now ? 1.0 : -1.0;
        for (int k = 0; k < 3; ++k) o[9 + k] = x[9 + k] / EULER_NORM;
        o[12] = x[13];

        double alt = x[1];
        double dx = x[0] - x[6], dy = x[1] - x[7], dz = x[2] - x[8];
        double loc = pow(pow(dx, 2.0) + pow(dy, 2.0) + pow(dz, 2.0), 0.5);
        int fired = x[14] > 0;
        double r = 0.0;
        r -= 0.0001 * loc;
        r -= t_angle * 10;
        if (alt < 2000) r -= 4;
        if (alt > 7000) r -= 4;
        int succ = 0;
        if (fired) {
            r -= 8;
            if (missile_prev && !locked_prev) succ = -1;
            else if (missile_prev && locked_prev) { succ = 1; fire_success = 1; }
        }
        if (x[13] <= 0.1 && fire_success) r += 600;
        reward[i] = r;
        int done = (alt < 500) || (alt > 10000);
        int ep = x[13] <= 0;
        if (ep) done = 1;
        int expert_fire = o[7] > 0 && o[8] > 0;   /* data/straight_line/ai_data_col.py:63-64 on the new observation */
        out_flags[i] = done | ep << 1 | fire_success << 2 | fired << 3 | expert_fire << 4 | ((succ & 0xFF) << 8);
    }
}

void env_oracle_offline(const double *s, const double *a, const double *s2, long n,
                        double *reward, int32_t *success, uint8_t *done, double *loc_out)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        const double *p = s + 13 * i, *act = a + 4 * i, *q = s2 + 13 * i;
        double loc = pow(pow(q[0] * 10000, 2.0) + pow(q[1] * 10000, 2.0) + pow(q[2] * 10000, 2.0), 0.5);
        double r = 0.0;
        r -= 0.0001 * loc;
        r -= q[6] * 10;
        int succ = 0;
        if (act[3] > 0) {
            r -= 8;
            if (p[8] > 0 && p[7] < 0) succ = -1;
            else if (p[8] > 0 && p[7] > 0) succ = 1;
        }
        if (q[12] < 0.1) r += 600;
        reward[i] = r;
        success[i] = succ;
        done[i] = q[12] <= 0.1;
        loc_out[i] = loc;
    }
}

void env_oracle_pow(const double *x, double e, long n, double *y)
{
    for (long i = 0; i < n; ++i) y[i] = pow(x[i], e);
}
