/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the integer/scan parts of the Diamond PPO hot path.
 *
 * Built by oracle/Makefile into oracle/_build/liboracle.so; loaded (ctypes) only by tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Never linked into the product library.
 * Pinned by tests/test_oracle.py against tests/golden/{gae,perm}.npz (outputs of the unmodified reference).
 *
 *   oracle_gae_f32          <- /root/reference/diamond/ppo.py:188-222 (calculate_advantage)
 *   oracle_mt19937_*        <- numpy 2.3.5 legacy RandomState: numpy/random/src/mt19937/mt19937.c
 *                              (init_genrand seeding, genrand), mtrand.pyx shuffle/permutation,
 *                              distributions.c random_interval -- what ppo.py:254 np.random.permutation runs.
 */
#include <stdint.h>
#include <stdlib.h>

/* ppo.py:201-220, float32, the reference's operation order (no fused multiply-add). */
void oracle_gae_f32(const float *rewards, const float *terminations, const float *truncations,
                    const float *values, const float *next_values, float *advantages, float *returns,
                    int T, int N, double gamma, double gae_lambda)
{
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);      /* python double product, then one cast */
    for (int e = 0; e < N; ++e) {
        volatile float adv = 0.0f;
        for (int t = T - 1; t >= 0; --t) {
            const long i = (long)t * N + e;
            volatile float nt = 1.0f - terminations[i];
            volatile float ntr = 1.0f - truncations[i];
            volatile float a = g * next_values[i];
            volatile float b = a * nt;
            volatile float c = rewards[i] + b;
            volatile float delta = c - values[i];
            volatile float k1 = gl * nt;
            volatile float k2 = k1 * ntr;
            volatile float k3 = k2 * adv;
            adv = delta + k3;
            advantages[i] = adv;
            if (returns) { volatile float rr = values[i] + adv; returns[i] = rr; }   /* ppo.py:241 */
        }
    }
}

typedef struct { uint32_t key[624]; int pos; } oracle_mt_t;

void oracle_mt19937_seed(oracle_mt_t *s, uint32_t seed)
{
    s->key[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->key[i] = 1812433253u * (s->key[i - 1] ^ (s->key[i - 1] >> 30)) + (uint32_t)i;
    s->pos = 624;
}

static void mt_gen(oracle_mt_t *s)
{
    uint32_t *k = s->key;
    for (int i = 0; i < 624; ++i) {
        uint32_t y = (k[i] & 0x80000000u) | (k[(i + 1) % 624] & 0x7fffffffu);
        k[i] = k[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    s->pos = 0;
}

uint32_t oracle_mt19937_next(oracle_mt_t *s)
{
    if (s->pos == 624) mt_gen(s);
    uint32_t y = s->key[s->pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* np.random.permutation(n): arange(n) shuffled from the end with masked-rejection draws. */
void oracle_permutation(oracle_mt_t *s, int64_t n, int64_t *out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = i;
    for (int64_t i = n - 1; i >= 1; --i) {
        uint64_t mask = (uint64_t)i;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16; mask |= mask >> 32;
        uint64_t j;
        do { j = oracle_mt19937_next(s) & mask; } while (j > (uint64_t)i);
        int64_t tmp = out[i]; out[i] = out[j]; out[j] = tmp;
    }
}

int oracle_mt_state_size(void) { return (int)sizeof(oracle_mt_t); }
