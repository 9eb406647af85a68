// host_rng.h -- exact, fast replicas of the two index draws the reference makes on the host per learn()
// call.  The reference owns the RNG streams (SURVEY 8b "RNG ownership"); these functions produce the SAME
// indices and leave the generator in the SAME state, only faster than the interpreter-level originals:
//   * HIRL.py:249 / BC.py:167  np.random.choice(N, B, replace=False) on the legacy global RandomState
//       = permutation(N)[:B] = Fisher-Yates from i=N-1 down to 1 with j = random_interval(i)
//         (masked rejection on 32-bit MT19937 outputs).  ~470 us in numpy for N=20000, ~100 us here.
//   * buffer.py:45  random.sample(memory, k) on CPython's global Random
//       = k draws of _randbelow(n) (top bits of one MT19937 output, redrawn while >= n) either from a shrinking pool
//         (small n) or with a "selected" set (large n): CPython Lib/random.py, Random.sample.
// Both run on a caller-provided copy of (or pointer to) the generator's MT19937 state {key[624], pos}; the Python
// side (utils/fast_rng.py) owns snapshot / compare / commit.  tests/test_host_logic.py checks indices AND final
// state against numpy / CPython themselves.
#pragma once
#include <cstdint>
#include <vector>

namespace hirl_host {

struct MT19937 {
    uint32_t *key;   // [624]
    int pos;
    inline void regenerate() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t MATRIX_A = 0x9908b0dfu, UPPER = 0x80000000u, LOWER = 0x7fffffffu;
        int kk;
        uint32_t y;
        for (kk = 0; kk < N - M; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + M] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        for (; kk < N - 1; ++kk) {
            y = (key[kk] & UPPER) | (key[kk + 1] & LOWER);
            key[kk] = key[kk + (M - N)] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        }
        y = (key[N - 1] & UPPER) | (key[0] & LOWER);
        key[N - 1] = key[M - 1] ^ (y >> 1) ^ (-(int32_t)(y & 1) & MATRIX_A);
        pos = 0;
    }
    static inline uint32_t temper(uint32_t y) {
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    inline uint32_t next() {
        if (pos >= 624) regenerate();
        return temper(key[pos++]);
    }
};

// numpy random_interval(max) for max <= 0xffffffff
static inline uint32_t random_interval(MT19937 &g, uint32_t max) {
    if (max == 0) return 0;
    uint32_t mask = max;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    while ((v = (g.next() & mask)) > max) {}
    return v;
}

}  // namespace hirl_host
