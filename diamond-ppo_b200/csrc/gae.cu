// GAE reverse-time scan (diamond/ppo.py:188-222) fused with returns (ppo.py:241) and the
// sum / sum-of-squares the advantage normalisation needs (ppo.py:243).
//
// Layout: all tensors time-major [T, N] f32, env contiguous.  One CTA owns 32 consecutive envs
// (one 128-byte line per time step and tensor) for the whole horizon: lane = env, warp = time
// chunk.  A_t = delta_t + c_t * A_{t+1} is a composition of affine maps, so every warp scans
// its own chunk of L steps assuming a zero carry, publishes the chunk's map (a0, P0) to shared
// memory, and after one barrier each thread folds the maps of the later chunks into its carry.
// All 5*L loads of a thread are independent and issued before the first use, so the whole CTA
// tile (5 * T * 128 B) is in flight at once -- the kernel is a pure HBM stream.
#include "common.cuh"

namespace {

constexpr int GAE_WARPS = 16;

template <int L>
__global__ void __launch_bounds__(GAE_WARPS * 32)
gae_kernel(const float* __restrict__ rewards, const float* __restrict__ terms, const float* __restrict__ truncs,
           const float* __restrict__ values, const float* __restrict__ next_values, float* __restrict__ adv_out,
           float* __restrict__ ret_out, double* __restrict__ stats, int T, int N, float gamma, float gamma_lambda)
{
    __shared__ float s_a0[GAE_WARPS][32];
    __shared__ float s_p0[GAE_WARPS][32];
    __shared__ float s_carry[32];
    __shared__ double s_red[2][GAE_WARPS];

    // Programmatic dependent launch: let the next kernel in the stream get its CTAs resident now
    // (it still waits at its own griddepcontrol.wait), and wait for everything this launch depends
    // on before the first global access.  Without the launch attribute both are no-ops.
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int env = blockIdx.x * 32 + lane;
    const bool env_ok = env < N;
    constexpr int SPAN = GAE_WARPS * L;                 // time steps covered per pass
    const int passes = (T + SPAN - 1) / SPAN;

    if (warp == 0) s_carry[lane] = 0.0f;
    float sum = 0.0f, sumsq = 0.0f;
    asm volatile("griddepcontrol.wait;" ::: "memory");

    for (int pass = passes - 1; pass >= 0; --pass) {
        const int t0 = pass * SPAN + warp * L;
        float r[L], te[L], tr[L], v[L], nv[L];
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int t = t0 + j;
            const bool ok = env_ok && t < T;
            const int64_t i = (int64_t)t * N + env;
            r[j] = ok ? __ldg(rewards + i) : 0.0f;
            te[j] = ok ? __ldg(terms + i) : 0.0f;
            tr[j] = ok ? __ldg(truncs + i) : 0.0f;
            v[j] = ok ? __ldg(values + i) : 0.0f;
            nv[j] = ok ? __ldg(next_values + i) : 0.0f;
        }
        // chunk-local scan with zero carry: a[j] = delta_j + c_j a[j+1], p[j] = c_j p[j+1]
        float a[L], p[L];
        float acc = 0.0f, prod = 1.0f;
#pragma unroll
        for (int j = L - 1; j >= 0; --j) {
            const float nt = 1.0f - te[j];
            const float ntr = 1.0f - tr[j];
            const float delta = r[j] + gamma * nv[j] * nt - v[j];      // ppo.py:206-210
            const float c = gamma_lambda * nt * ntr;                    // ppo.py:215-218
            acc = delta + c * acc;
            prod = c * prod;
            a[j] = acc;
            p[j] = prod;
        }
        s_a0[warp][lane] = a[0];
        s_p0[warp][lane] = p[0];
        __syncthreads();
        // carry entering this chunk = true advantage at the first step of the next chunk
        float carry = s_carry[lane];
#pragma unroll
        for (int w = GAE_WARPS - 1; w >= 1; --w)
            if (w > warp) carry = s_a0[w][lane] + s_p0[w][lane] * carry;
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int t = t0 + j;
            if (env_ok && t < T) {
                const int64_t i = (int64_t)t * N + env;
                const float A = a[j] + p[j] * carry;
                adv_out[i] = A;
                if (ret_out) ret_out[i] = v[j] + A;                     // ppo.py:241
                sum += A;
                sumsq += A * A;
            }
        }
        if (passes > 1) {
            __syncthreads();
            if (warp == 0) s_carry[lane] = a[0] + p[0] * carry;
            __syncthreads();
        }
    }

    if (stats) {
        double ds = warp_sum_d((double)sum), dq = warp_sum_d((double)sumsq);
        if (lane == 0) { s_red[0][warp] = ds; s_red[1][warp] = dq; }
        __syncthreads();
        if (warp == 0) {
            ds = lane < GAE_WARPS ? s_red[0][lane] : 0.0;
            dq = lane < GAE_WARPS ? s_red[1][lane] : 0.0;
            ds = warp_sum_d(ds); dq = warp_sum_d(dq);
            if (lane == 0) { atomicAdd(stats, ds); atomicAdd(stats + 1, dq); }
        }
    }
}

__global__ void adv_normalize_kernel(const float* __restrict__ adv, float* __restrict__ out,
                                     const double* __restrict__ stats, int64_t count, int64_t n)
{
    const double mean = stats[0] / (double)count;
    double var = (stats[1] - stats[0] * mean) / (double)(count - 1);      // unbiased (torch.std default)
    var = var > 0.0 ? var : 0.0;
    const float meanf = (float)mean;
    const float denom = (float)sqrt(var) + 1e-6f;                          // ppo.py:243
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (adv[i] - meanf) / denom;
}

}  // namespace

extern "C" int dppo_gae_f32(dppo_ctx* ctx, const float* rewards, const float* terminations, const float* truncations,
                            const float* values, const float* next_values, float* advantages, float* returns,
                            double* stats, int T, int N, double gamma, double gae_lambda, void* stream)
{
    if (!ctx) return 1;
    if (T <= 0 || N <= 0) DPPO_FAIL(ctx, "dppo_gae_f32: empty shape T=%d N=%d", T, N);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = (N + 31) / 32;
    const float g = (float)gamma, gl = (float)(gamma * gae_lambda);
    const int per_warp = (T + GAE_WARPS - 1) / GAE_WARPS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks);
    cfg.blockDim = dim3(GAE_WARPS * 32);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define GAE_LAUNCH(L) cudaLaunchKernelEx(&cfg, gae_kernel<L>, rewards, terminations, truncations, values, \
                                         next_values, advantages, returns, stats, T, N, g, gl)
    if (per_warp <= 1) GAE_LAUNCH(1);
    else if (per_warp <= 2) GAE_LAUNCH(2);
    else if (per_warp <= 4) GAE_LAUNCH(4);
    else GAE_LAUNCH(8);                       // T > 128 runs several passes of 128 steps
#undef GAE_LAUNCH
    DPPO_CHECK_LAUNCH(ctx, "gae_kernel");
    return 0;
}

extern "C" int dppo_adv_normalize_f32(dppo_ctx* ctx, const float* adv, float* out, const double* stats, int64_t count,
                                      int64_t n, void* stream)
{
    if (!ctx) return 1;
    if (count < 2) DPPO_FAIL(ctx, "dppo_adv_normalize_f32: need at least 2 samples (std is unbiased), got %lld", (long long)count);
    if (n <= 0) return 0;
    int blocks = (int)((n + 255) / 256);
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    adv_normalize_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(adv, out, stats, count, n);
    DPPO_CHECK_LAUNCH(ctx, "adv_normalize_kernel");
    return 0;
}
