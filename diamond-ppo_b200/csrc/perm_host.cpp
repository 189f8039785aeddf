// Host side of the minibatch permutation (diamond/ppo.py:252-255).
//
// The reference draws np.random.permutation(B) from numpy's legacy global RandomState: MT19937 + a Fisher-Yates shuffle from
// the end whose bounded draws are masked-rejection 32-bit draws (numpy/random/_legacy: legacy shuffle -> rk_interval).  It is a
// strictly sequential algorithm and must be reproduced bit for bit, but at 524 288 indices it is also the longest host-side
// step of a learn() -- under data parallelism EVERY rank needs the same global permutation while its GPU finishes an epoch in
// a fraction of the single-GPU time.  So the work is split into three passes that a modern core can stream:
//   1. raw MT19937 words in blocks of 624 (the twist has a dependency distance of 227 words: it vectorises) + tempering;
//   2. the rejection filter, branch-free: every masked candidate is stored, the cursor only advances on acceptance
//      (the data-dependent branch of the textbook loop mispredicts on ~27 % of the draws and serialises the memory accesses);
//   3. the swaps, driven by the accepted draws with a software prefetch of the random partner a few iterations ahead.
// Works on the caller's copy of the MT19937 state (np.random.get_state() / set_state()), emits int32 indices ready for upload,
// and is meant to run on a worker thread while the GPU is busy with the previous epoch.  `dppo_permutation_mt19937_skip`
// runs passes 1-2 only: it advances the stream exactly as a permutation of n would (ranks that do not own an epoch's
// permutation under data parallelism).
#include <stdint.h>

#include <vector>

#include "dppo.h"

#if defined(__GNUC__) && defined(__x86_64__) && !defined(__CUDACC__)
#define DPPO_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define DPPO_CLONES
#endif

namespace {
constexpr int MT_N = 624, MT_M = 397;

DPPO_CLONES void mt_refill(uint32_t* k)
{
    constexpr uint32_t UP = 0x80000000u, LO = 0x7fffffffu, MAT = 0x9908b0dfu;
    int i = 0;
    for (; i < MT_N - MT_M; ++i) {
        const uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
        k[i] = k[i + MT_M] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
    }
    for (; i < MT_N - 1; ++i) {
        const uint32_t y = (k[i] & UP) | (k[i + 1] & LO);
        k[i] = k[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
    }
    const uint32_t y = (k[MT_N - 1] & UP) | (k[0] & LO);
    k[MT_N - 1] = k[MT_M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & MAT);
}

DPPO_CLONES void temper_block(const uint32_t* __restrict__ k, uint32_t* __restrict__ t, int from)
{
    for (int i = from; i < MT_N; ++i) {
        uint32_t y = k[i];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        t[i] = y;
    }
}

// Passes 1 + 2: jl[c] (optional) = the accepted draw j of Fisher-Yates step i = n-1-c, c = 0 .. n-2.  Returns the stream position.
int draw_sequence(uint32_t* key, int pos, int64_t n, int32_t* jl)
{
    uint32_t tw[MT_N];
    int cur = pos;                               // next unread word of the current block (624: block exhausted)
    if (cur < MT_N) temper_block(key, tw, cur);
    int64_t i = n - 1, cnt = 0;
    while (i >= 1) {
        uint32_t mask = (uint32_t)i;             // rk_interval: smallest 2^k - 1 >= i
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        const int64_t lo = (int64_t)(mask >> 1); // the mask stays valid while i > lo
        while (i > lo) {
            if (cur == MT_N) { mt_refill(key); temper_block(key, tw, 0); cur = 0; }
            int w = cur;
            if (jl) {
                for (; w < MT_N && i > lo; ++w) {
                    const uint32_t cand = tw[w] & mask;
                    const int64_t ok = cand <= (uint32_t)i;
                    jl[cnt] = (int32_t)cand;     // overwritten by the next candidate unless accepted
                    cnt += ok;
                    i -= ok;
                }
            } else {
                for (; w < MT_N && i > lo; ++w) i -= (tw[w] & mask) <= (uint32_t)i;
            }
            cur = w;
        }
    }
    return cur;
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
    const int pos = *pos_io;
    if (pos < 0 || pos > 624) return 1;
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)i;
    if (n < 2) return 0;
    static thread_local std::vector<int32_t> scratch;
    if ((int64_t)scratch.size() < n + 64) scratch.resize((size_t)n + 64);
    int32_t* jl = scratch.data();
    *pos_io = draw_sequence(key, pos, n, jl);
    for (int64_t c = n - 1; c < n + 63; ++c) jl[c] = 0;          // prefetch targets past the end
    constexpr int PF = 24;
    for (int64_t c = 0; c < n - 1; ++c) {
        const int64_t i = n - 1 - c;
        const int32_t j = jl[c];
        __builtin_prefetch(out + jl[c + PF], 1, 1);
        const int32_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
    return 0;
}

extern "C" int dppo_permutation_mt19937_skip(uint32_t* key, int32_t* pos_io, int64_t n)
{
    if (!key || !pos_io || n < 0 || n > 0x7fffffffLL) return 1;
    const int pos = *pos_io;
    if (pos < 0 || pos > 624) return 1;
    if (n < 2) return 0;
    *pos_io = draw_sequence(key, pos, n, nullptr);
    return 0;
}
