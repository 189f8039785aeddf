// Device side of the minibatch permutation (diamond/ppo.py:252-255).
//
//   perm_feistel_kernel      : the FAST (non-parity) generator of SURVEY.md 2.2 K4a.  np.random.permutation is a strictly
//                              sequential algorithm (MT19937 + Fisher-Yates, perm_host.cpp keeps it bit-exact on a host thread);
//                              this kernel instead evaluates a keyed bijection of [0, n) per element -- an unbalanced Feistel network
//                              over the next power of two with cycle walking -- so a permutation costs one 4-byte store per index,
//                              needs no sort and no host work, and is a pure function of (seed, counter): every data-parallel
//                              rank evaluates the same permutation without communication.
//   perm_shard_filter_kernel : env-sharded data parallelism.  Every rank holds the SAME global permutation of the
//                              concatenated buffer (flat index t*N_global + env); rank r owns envs [lo, lo + n_local).  One block
//                              per global minibatch keeps, in permutation order (stable compaction), the members that fall into the
//                              rank's shard, re-indexed to the rank-local buffer (t*n_local + env - lo), and pads the list with -1
//                              up to a fixed M_pad rows: the update kernels then run with the same shapes on every step (CUDA-graph
//                              replay) and treat negative indices as rows that contribute nothing.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x, uint32_t k)
{
    x ^= k;
    x *= 0x9E3779B1u; x ^= x >> 15;
    x *= 0x85EBCA77u; x ^= x >> 13;
    x *= 0xC2B2AE3Du; x ^= x >> 16;
    return x;
}

constexpr int FEISTEL_ROUNDS = 8;        // pairs of half-rounds

struct FeistelKeys { uint32_t k[2 * FEISTEL_ROUNDS]; };

// Bijection of [0, 2^(lb+rb)): alternately L ^= F(R) and R ^= F(L) -- each step is invertible whatever F is, and the two halves
// may have different widths (odd bit counts need no special case).
__device__ __forceinline__ uint32_t feistel(uint32_t x, int lb, int rb, const FeistelKeys& key)
{
    const uint32_t mask_l = (1u << lb) - 1u, mask_r = (1u << rb) - 1u;
    uint32_t L = x >> rb, R = x & mask_r;
#pragma unroll
    for (int r = 0; r < FEISTEL_ROUNDS; ++r) {
        L ^= mix32(R, key.k[2 * r]) & mask_l;
        R ^= mix32(L, key.k[2 * r + 1]) & mask_r;
    }
    return (L << rb) | R;
}

__global__ void __launch_bounds__(256)
perm_feistel_kernel(FeistelKeys key, int bits, uint32_t n, int32_t* __restrict__ out)
{
    const int lb = bits / 2, rb = bits - lb;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t x = i;
        do { x = feistel(x, lb, rb, key); } while (x >= n);      // cycle walking: the restriction of a bijection to [0, n)
        out[i] = (int32_t)x;
    }
}

constexpr int FILTER_THREADS = 1024;

__global__ void __launch_bounds__(FILTER_THREADS)
perm_shard_filter_kernel(const int32_t* __restrict__ perm, int64_t M, int n_global, int lo, int n_local, int64_t M_pad,
                         int32_t* __restrict__ idx_out, int32_t* __restrict__ counts, int32_t* __restrict__ overflow)
{
    __shared__ int s_warp[FILTER_THREADS / 32];
    __shared__ int s_base;
    const int k = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t* src = perm + (int64_t)k * M;
    int32_t* dst = idx_out + (int64_t)k * M_pad;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int64_t c = 0; c < M; c += FILTER_THREADS) {
        const int64_t i = c + threadIdx.x;
        int local = -1;
        if (i < M) {
            const int p = __ldg(src + i);
            const int t = p / n_global, env = p - t * n_global;
            if (env >= lo && env < lo + n_local) local = t * n_local + (env - lo);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, local >= 0);
        if (lane == 0) s_warp[warp] = __popc(ballot);
        __syncthreads();
        int before = s_base;
        for (int w = 0; w < warp; ++w) before += s_warp[w];
        const int pos = before + __popc(ballot & ((1u << lane) - 1u));
        if (local >= 0 && pos < M_pad) dst[pos] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < FILTER_THREADS / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int count = s_base;
    for (int64_t j = (int64_t)count + threadIdx.x; j < M_pad; j += FILTER_THREADS) dst[j] = -1;
    if (threadIdx.x == 0) {
        counts[k] = count < M_pad ? count : (int)M_pad;
        if (count > M_pad) atomicOr(overflow, 1);
    }
}

// SplitMix64: key schedule of the Feistel network from (seed, counter) on the host
inline uint64_t splitmix64(uint64_t& s)
{
    s += 0x9E3779B97F4A7C15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace

extern "C" int dppo_permutation_device(dppo_ctx* ctx, uint64_t seed, uint64_t counter, int64_t n, int32_t* out, void* stream)
{
    if (!ctx) return 1;
    if (!out || n < 1 || n > 0x7fffffffLL) DPPO_FAIL(ctx, "permutation_device: n must be in [1, 2^31)");
    int bits = 2;                                   // at least one bit per half
    while (((int64_t)1 << bits) < n) ++bits;
    FeistelKeys key;
    uint64_t s = seed ^ (counter * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    for (int i = 0; i < 2 * FEISTEL_ROUNDS; i += 2) {
        const uint64_t z = splitmix64(s);
        key.k[i] = (uint32_t)z; key.k[i + 1] = (uint32_t)(z >> 32);
    }
    int blocks = (int)((n + 255) / 256);
    if (blocks > 8 * ctx->sm_count) blocks = 8 * ctx->sm_count;
    perm_feistel_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(key, bits, (uint32_t)n, out);
    DPPO_CHECK_LAUNCH(ctx, "perm_feistel_kernel");
    return 0;
}

extern "C" int dppo_perm_shard_filter(dppo_ctx* ctx, const int32_t* perm, int64_t B_global, int n_global_envs, int env_lo, int n_local,
                                      int num_minibatches, int64_t M_pad, int32_t* idx_out, int32_t* counts, int32_t* overflow,
                                      void* stream)
{
    if (!ctx) return 1;
    if (!perm || !idx_out || !counts || !overflow) DPPO_FAIL(ctx, "perm_shard_filter: null argument");
    if (num_minibatches < 1 || B_global % num_minibatches != 0 || n_global_envs < 1 || n_local < 1 || env_lo < 0 ||
        env_lo + n_local > n_global_envs || M_pad < 1 || B_global % n_global_envs != 0)
        DPPO_FAIL(ctx, "perm_shard_filter: bad shape");
    perm_shard_filter_kernel<<<num_minibatches, FILTER_THREADS, 0, (cudaStream_t)stream>>>(perm, B_global / num_minibatches, n_global_envs,
                                                                                          env_lo, n_local, M_pad, idx_out, counts, overflow);
    DPPO_CHECK_LAUNCH(ctx, "perm_shard_filter_kernel");
    return 0;
}
