// Host side of the minibatch permutation (diamond/ppo.py:252-255).
//
// The reference draws np.random.permutation(B) from numpy's legacy global RandomState, which is a
// strictly sequential algorithm (MT19937 + Fisher-Yates with masked-rejection 32-bit draws) and
// must be reproduced bit for bit.  This is a tight C++ version that works on the caller's copy of
// the MT19937 state (np.random.get_state() / set_state()), emits int32 indices ready for upload,
// and is meant to run on a worker thread while the GPU is busy with the previous epoch.
#include <stdint.h>

#include "dppo.h"

namespace {
inline void mt_refill(uint32_t* k)
{
    constexpr int N = 624, M = 397;
    constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAT = 0x9908b0dfu;
    int i = 0;
    for (; i < N - M; ++i) {
        const uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
        k[i] = k[i + M] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    for (; i < N - 1; ++i) {
        const uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
        k[i] = k[i + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
    }
    const uint32_t y = (k[N - 1] & UP) | (k[0] & LO);
    k[N - 1] = k[M - 1] ^ (y >> 1) ^ ((y & 1u) ? MAT : 0u);
}
inline uint32_t temper(uint32_t y)
{
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
}  // namespace

extern "C" int dppo_mt19937_seed(uint32_t* key, int32_t* pos, uint32_t seed)
{
    if (!key || !pos) return 1;
    key[0] = seed;
    for (uint32_t i = 1; i < 624; ++i) key[i] = 1812433253u * (key[i - 1] ^ (key[i - 1] >> 30)) + i;
    *pos = 624;
    return 0;
}

extern "C" int dppo_permutation_mt19937(uint32_t* key, int32_t* pos_io, int64_t n, int32_t* out)
{
    if (!key || !pos_io || !out || n < 0 || n > 0x7fffffffLL) return 1;
    int pos = *pos_io;
    if (pos < 0 || pos > 624) return 1;
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)i;
    for (int64_t i = n - 1; i >= 1; --i) {
        uint32_t mask = (uint32_t)i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t j;
        do {
            if (pos == 624) { mt_refill(key); pos = 0; }
            j = temper(key[pos++]) & mask;
        } while (j > (uint32_t)i);
        const int32_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
    *pos_io = pos;
    return 0;
}
