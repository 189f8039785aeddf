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
//
// Two variants share the scan: `gae_tma_kernel` stages the five [T x envs] input tiles through shared
// memory with one 2-D TMA bulk-tensor load each (cp.async.bulk.tensor, completion on an mbarrier), so the
// whole tile is in flight without occupying L1 miss slots -- the register variant `gae_kernel` keeps
// 5*L loads per thread in flight and was measured L1-miss-slot bound (3.2 TB/s at 4096 x 128).  The
// TMA variant also sizes the env block so that the grid covers all SMs (28 envs x 147 CTAs at N=4096).
// The register variant remains for shapes TMA cannot describe (N not a multiple of 4, unaligned bases).
#include <cuda.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int GAE_WARPS = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "GAE_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra GAE_DONE_%=;\n\t"
        "bra GAE_WAIT_%=;\n\t"
        "GAE_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

// One affine-map scan shared by both kernels (see the header comment).
template <int L>
struct GaeScan {
    float a[L], p[L];
    __device__ __forceinline__ void local(const float (&r)[L], const float (&te)[L], const float (&tr)[L], const float (&v)[L],
                                          const float (&nv)[L], float gamma, float gamma_lambda)
    {
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
    }
};

template <int L>
__global__ void __launch_bounds__(GAE_WARPS * 32)
gae_tma_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_te,
               const __grid_constant__ CUtensorMap tm_tr, const __grid_constant__ CUtensorMap tm_v,
               const __grid_constant__ CUtensorMap tm_nv, float* __restrict__ adv_out, float* __restrict__ ret_out,
               double* __restrict__ stats, int T, int N, int epc, int box_rows, int region_bytes, float gamma, float gamma_lambda)
{
    extern __shared__ __align__(128) unsigned char dyn[];
    __shared__ __align__(8) uint64_t mbar, mbar_b;
    __shared__ float s_a0[GAE_WARPS][32];
    __shared__ float s_p0[GAE_WARPS][32];
    __shared__ float s_carry[32];
    __shared__ double s_red[2][GAE_WARPS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int env0 = blockIdx.x * epc;
    const int env = env0 + lane;
    const bool env_ok = lane < epc && env < N;
    constexpr int SPAN = GAE_WARPS * L;
    const int passes = (T + SPAN - 1) / SPAN;
    if (threadIdx.x == 0) { mbar_init(&mbar, 1); mbar_init(&mbar_b, 1); }
    if (warp == 0) s_carry[lane] = 0.0f;
    __syncthreads();

    const float* s_r = reinterpret_cast<const float*>(dyn);
    const float* s_te = reinterpret_cast<const float*>(dyn + region_bytes);
    const float* s_tr = reinterpret_cast<const float*>(dyn + 2 * region_bytes);
    const float* s_v = reinterpret_cast<const float*>(dyn + 3 * region_bytes);
    const float* s_nv = reinterpret_cast<const float*>(dyn + 4 * region_bytes);
    float sum = 0.0f, sumsq = 0.0f;
    uint32_t parity = 0;

    for (int pass = passes - 1; pass >= 0; --pass) {
        const int tbase = pass * SPAN;
        // Programmatic dependent launch: rewards and the two masks are rollout data that no preceding kernel of the stream
        // writes, so their tiles are requested before griddepcontrol.wait and stream in while the predecessor (the
        // pre-update pass that produces values / next_values, or the previous GAE launch) drains; the dependents of this
        // grid are released at once so that the same overlap continues down the stream.
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&mbar, 3u * (uint32_t)(box_rows * epc * 4));
            tma_load_2d(dyn, &tm_r, env0, tbase, &mbar);
            tma_load_2d(dyn + region_bytes, &tm_te, env0, tbase, &mbar);
            tma_load_2d(dyn + 2 * region_bytes, &tm_tr, env0, tbase, &mbar);
        }
        if (pass == passes - 1) {
            asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            asm volatile("griddepcontrol.wait;" ::: "memory");
        }
        if (threadIdx.x == 0) {
            mbar_expect_tx(&mbar_b, 2u * (uint32_t)(box_rows * epc * 4));
            tma_load_2d(dyn + 3 * region_bytes, &tm_v, env0, tbase, &mbar_b);
            tma_load_2d(dyn + 4 * region_bytes, &tm_nv, env0, tbase, &mbar_b);
        }
        mbar_wait(&mbar, parity);
        mbar_wait(&mbar_b, parity);
        parity ^= 1u;
        float r[L], te[L], tr[L], v[L], nv[L];
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int row = warp * L + j;
            const bool ok = row < box_rows && lane < epc;      // rows past T inside the box are zero-filled by TMA
            const int i = row * epc + lane;
            r[j] = ok ? s_r[i] : 0.0f;
            te[j] = ok ? s_te[i] : 0.0f;
            tr[j] = ok ? s_tr[i] : 0.0f;
            v[j] = ok ? s_v[i] : 0.0f;
            nv[j] = ok ? s_nv[i] : 0.0f;
        }
        GaeScan<L> sc;
        sc.local(r, te, tr, v, nv, gamma, gamma_lambda);
        s_a0[warp][lane] = sc.a[0];
        s_p0[warp][lane] = sc.p[0];
        __syncthreads();
        float carry = s_carry[lane];
#pragma unroll
        for (int w = GAE_WARPS - 1; w >= 1; --w)
            if (w > warp) carry = s_a0[w][lane] + s_p0[w][lane] * carry;
#pragma unroll
        for (int j = 0; j < L; ++j) {
            const int t = tbase + warp * L + j;
            if (env_ok && t < T) {
                const int64_t i = (int64_t)t * N + env;
                const float A = sc.a[j] + sc.p[j] * carry;
                adv_out[i] = A;
                if (ret_out) ret_out[i] = v[j] + A;                     // ppo.py:241
                sum += A;
                sumsq += A * A;
            }
        }
        if (passes > 1) {
            __syncthreads();
            if (warp == 0) s_carry[lane] = sc.a[0] + sc.p[0] * carry;
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

// Pipelined variant for horizons of at least 128 steps: the 128 x envs tile of a pass is requested as four 32-step chunks,
// latest steps first, each with its own mbarrier.  The four warps of a chunk scan and STORE as soon as their chunk has
// landed -- the reverse-time recurrence only needs the (already finished) later chunk's carry -- so the output stream
// overlaps the rest of the input stream instead of following it.
//   settled != 0: the caller guarantees that none of the five inputs is written by the kernel immediately preceding this
//   launch in the stream (PPO.learn(): that kernel is the 16-byte statistics fill).  All input tiles are then requested
//   before griddepcontrol.wait, which is only executed before the first global write; back-to-back launches overlap
//   their load and store phases.  settled == 0: only rewards and the masks (rollout data) are requested early.
constexpr int GP_CHUNKS = 4, GP_L = 8, GP_ROWS = 32, GP_WPC = GAE_WARPS / GP_CHUNKS;      // 4 warps x 8 steps per chunk

__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__global__ void __launch_bounds__(GAE_WARPS * 32)
gae_pipe_kernel(const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_te,
                const __grid_constant__ CUtensorMap tm_tr, const __grid_constant__ CUtensorMap tm_v,
                const __grid_constant__ CUtensorMap tm_nv, float* __restrict__ adv_out, float* __restrict__ ret_out,
                double* __restrict__ stats, int T, int N, int epc, int chunk_bytes, float gamma, float gamma_lambda, int settled)
{
    extern __shared__ __align__(128) unsigned char dyn[];       // [tensor 0..4][chunk 0..3][32 rows x epc]
    __shared__ __align__(8) uint64_t cbar[GP_CHUNKS];
    __shared__ float s_a0[GAE_WARPS][32];
    __shared__ float s_p0[GAE_WARPS][32];
    __shared__ float s_ccarry[GP_CHUNKS + 1][32];               // [c]: advantage at the first step of chunk c (= carry into chunk c-1);
                                                                // [GP_CHUNKS]: carry entering the pass (0, or the later pass's first step)
    __shared__ double s_red[2][GAE_WARPS];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int chunk = warp / GP_WPC, wic = warp % GP_WPC;       // chunk of this warp, warp index inside the chunk
    const int env0 = blockIdx.x * epc;
    const int env = env0 + lane;
    const bool env_ok = lane < epc && env < N;
    constexpr int SPAN = GAE_WARPS * GP_L;                      // 128 steps per pass
    const int passes = (T + SPAN - 1) / SPAN;
    const int region = GP_CHUNKS * chunk_bytes;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int c = 0; c < GP_CHUNKS; ++c) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&cbar[c])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) s_ccarry[GP_CHUNKS][lane] = 0.0f;
    __syncthreads();

    float sum = 0.0f, sumsq = 0.0f;
    uint32_t parity = 0;
    bool waited = false;
    for (int pass = passes - 1; pass >= 0; --pass) {
        const int tbase = pass * SPAN;
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int c = GP_CHUNKS - 1; c >= 0; --c) {
                mbar_expect_tx(&cbar[c], 5u * (uint32_t)(GP_ROWS * epc * 4));
                tma_load_2d(dyn + 0 * region + c * chunk_bytes, &tm_r, env0, tbase + c * GP_ROWS, &cbar[c]);
                tma_load_2d(dyn + 1 * region + c * chunk_bytes, &tm_te, env0, tbase + c * GP_ROWS, &cbar[c]);
                tma_load_2d(dyn + 2 * region + c * chunk_bytes, &tm_tr, env0, tbase + c * GP_ROWS, &cbar[c]);
                if (settled) {
                    tma_load_2d(dyn + 3 * region + c * chunk_bytes, &tm_v, env0, tbase + c * GP_ROWS, &cbar[c]);
                    tma_load_2d(dyn + 4 * region + c * chunk_bytes, &tm_nv, env0, tbase + c * GP_ROWS, &cbar[c]);
                }
            }
        }
        if (pass == passes - 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (!settled) {
            if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
            if (threadIdx.x == 0) {
                for (int c = GP_CHUNKS - 1; c >= 0; --c) {
                    tma_load_2d(dyn + 3 * region + c * chunk_bytes, &tm_v, env0, tbase + c * GP_ROWS, &cbar[c]);
                    tma_load_2d(dyn + 4 * region + c * chunk_bytes, &tm_nv, env0, tbase + c * GP_ROWS, &cbar[c]);
                }
            }
        }
        mbar_wait(&cbar[chunk], parity);
        parity ^= 1u;
        const float* base = reinterpret_cast<const float*>(dyn + chunk * chunk_bytes);
        const int fl = region / 4;                                  // floats between tensors
        float r[GP_L], te[GP_L], tr[GP_L], v[GP_L], nv[GP_L];
#pragma unroll
        for (int j = 0; j < GP_L; ++j) {
            const int i = (wic * GP_L + j) * epc + lane;            // rows past T inside the box are zero-filled by TMA
            const bool ok = lane < epc;
            r[j] = ok ? base[i] : 0.0f;
            te[j] = ok ? base[fl + i] : 0.0f;
            tr[j] = ok ? base[2 * fl + i] : 0.0f;
            v[j] = ok ? base[3 * fl + i] : 0.0f;
            nv[j] = ok ? base[4 * fl + i] : 0.0f;
        }
        GaeScan<GP_L> sc;
        sc.local(r, te, tr, v, nv, gamma, gamma_lambda);
        s_a0[warp][lane] = sc.a[0];
        s_p0[warp][lane] = sc.p[0];
        // the chunk's four warps have published their maps; the later chunk (or the previous pass) has published its carry
        named_bar_sync(1 + chunk, GP_WPC * 32);
        if (chunk < GP_CHUNKS - 1) named_bar_sync(1 + GP_CHUNKS + chunk, GP_WPC * 32 + 32);   // released by the first warp of chunk + 1
        float carry = s_ccarry[chunk + 1][lane];
#pragma unroll
        for (int w = GP_WPC - 1; w >= 1; --w)
            if (w > wic) carry = s_a0[chunk * GP_WPC + w][lane] + s_p0[chunk * GP_WPC + w][lane] * carry;
        if (wic == 0) {
            // advantage at the first step of this chunk = carry entering the chunk before it in time
            s_ccarry[chunk][lane] = sc.a[0] + sc.p[0] * carry;
            if (chunk > 0) {
                __threadfence_block();
                named_bar_arrive(1 + GP_CHUNKS + chunk - 1, GP_WPC * 32 + 32);
            }
        }
        if (settled && !waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
#pragma unroll
        for (int j = 0; j < GP_L; ++j) {
            const int t = tbase + warp * GP_L + j;
            if (env_ok && t < T) {
                const int64_t i = (int64_t)t * N + env;
                const float A = sc.a[j] + sc.p[j] * carry;
                adv_out[i] = A;
                if (ret_out) ret_out[i] = v[j] + A;                     // ppo.py:241
                sum += A;
                sumsq += A * A;
            }
        }
        if (passes > 1) {
            __syncthreads();
            if (warp == 0) s_ccarry[GP_CHUNKS][lane] = s_ccarry[0][lane];
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

// ---- host: tensor maps ----------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct TmEntry { const void* ptr; int T, N, epc, rows; CUtensorMap map; };
struct TmCache { std::vector<TmEntry> e; size_t next = 0; };
void tm_cache_free(void* p) { delete static_cast<TmCache*>(p); }

bool get_tensor_map(dppo_ctx* ctx, const float* ptr, int T, int N, int epc, int rows, CUtensorMap* out)
{
    if (!ctx->tm_cache) { ctx->tm_cache = new TmCache(); ctx->tm_cache_free = tm_cache_free; }
    TmCache* c = static_cast<TmCache*>(ctx->tm_cache);
    for (const TmEntry& t : c->e)
        if (t.ptr == ptr && t.T == T && t.N == N && t.epc == epc && t.rows == rows) { *out = t.map; return true; }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    TmEntry t;
    t.ptr = ptr; t.T = T; t.N = N; t.epc = epc; t.rows = rows;
    const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)T};
    const cuuint64_t gstride[1] = {(cuuint64_t)N * 4};
    const cuuint32_t box[2] = {(cuuint32_t)epc, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&t.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (c->e.size() < 256) c->e.push_back(t);
    else { c->e[c->next] = t; c->next = (c->next + 1) % 256; }
    *out = t.map;
    return true;
}

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

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int env = blockIdx.x * 32 + lane;
    const bool env_ok = env < N;
    constexpr int SPAN = GAE_WARPS * L;                 // time steps covered per pass
    const int passes = (T + SPAN - 1) / SPAN;

    if (warp == 0) s_carry[lane] = 0.0f;
    float sum = 0.0f, sumsq = 0.0f;

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

    // TMA variant: needs 16-byte row pitch and base alignment (tensor-map requirements)
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (ctx->gae_variant != 1 && N % 4 == 0 && al16(rewards) && al16(terminations) && al16(truncations) && al16(values) && al16(next_values)) {
        int epc = (N + ctx->sm_count - 1) / ctx->sm_count;
        epc = (epc + 3) / 4 * 4;
        if (epc > 32) epc = 32;
        if (T >= GAE_WARPS * GP_L && ctx->gae_variant != 2) {
            // pipelined kernel: 32-step chunks with their own barriers, stores overlap the remaining loads
            const int chunk_bytes = (GP_ROWS * epc * 4 + 127) / 128 * 128;
            CUtensorMap m[5];
            const float* ptrs[5] = {rewards, terminations, truncations, values, next_values};
            bool ok = true;
            for (int i = 0; i < 5 && ok; ++i) ok = get_tensor_map(ctx, ptrs[i], T, N, epc, GP_ROWS, &m[i]);
            if (ok) {
                const int grid = (N + epc - 1) / epc;
                const size_t smem = (size_t)5 * GP_CHUNKS * chunk_bytes;
                cudaLaunchConfig_t lc = {};
                lc.gridDim = dim3(grid); lc.blockDim = dim3(GAE_WARPS * 32); lc.dynamicSmemBytes = smem; lc.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                attr[0].val.programmaticStreamSerializationAllowed = 1;
                // level 0 (public API default): no programmatic dependent launch -- the kernel starts after its predecessor has
                // completed and nothing is loaded early; 1: rollout tensors early; 2: all five inputs early (dppo.h)
                lc.attrs = attr; lc.numAttrs = ctx->gae_inputs_settled >= 1 ? 1 : 0;
                cudaFuncSetAttribute(gae_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaLaunchKernelEx(&lc, gae_pipe_kernel, m[0], m[1], m[2], m[3], m[4], advantages, returns, stats, T, N, epc, chunk_bytes,
                                   g, gl, ctx->gae_inputs_settled >= 2 ? 1 : 0);
                DPPO_CHECK_LAUNCH(ctx, "gae_pipe_kernel");
                return 0;
            }
        }
        const int L = per_warp <= 1 ? 1 : per_warp <= 2 ? 2 : per_warp <= 4 ? 4 : 8;
        const int span = GAE_WARPS * L;
        const int rows = T < span ? T : span;
        const int region = (rows * epc * 4 + 127) / 128 * 128;
        CUtensorMap m[5];
        const float* ptrs[5] = {rewards, terminations, truncations, values, next_values};
        bool ok = true;
        for (int i = 0; i < 5 && ok; ++i) ok = get_tensor_map(ctx, ptrs[i], T, N, epc, rows, &m[i]);
        if (ok) {
            const int grid = (N + epc - 1) / epc;
            const size_t smem = (size_t)5 * region;
            cudaLaunchConfig_t lc = {};
            lc.gridDim = dim3(grid); lc.blockDim = dim3(GAE_WARPS * 32); lc.dynamicSmemBytes = smem; lc.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = attr; lc.numAttrs = ctx->gae_inputs_settled >= 1 ? 1 : 0;
#define GAE_TMA_LAUNCH(LL)                                                                                              \
            do {                                                                                                        \
                cudaFuncSetAttribute(gae_tma_kernel<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
                cudaLaunchKernelEx(&lc, gae_tma_kernel<LL>, m[0], m[1], m[2], m[3], m[4], advantages, returns, stats,   \
                                   T, N, epc, rows, region, g, gl);                                                     \
            } while (0)
            if (L == 1) GAE_TMA_LAUNCH(1); else if (L == 2) GAE_TMA_LAUNCH(2); else if (L == 4) GAE_TMA_LAUNCH(4); else GAE_TMA_LAUNCH(8);
#undef GAE_TMA_LAUNCH
            DPPO_CHECK_LAUNCH(ctx, "gae_tma_kernel");
            return 0;
        }
    }
#define GAE_LAUNCH(L) gae_kernel<L><<<blocks, GAE_WARPS * 32, 0, st>>>(rewards, terminations, truncations, values, \
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
