// Env-sharded data parallelism: the per-minibatch exchange step.
//
// Every rank (one process per GPU) computes the gradient of its shard of the global minibatch; the reference-equivalent
// update needs their SUM before clip_grad_norm_ (ppo.py:284) and Adam (ppo.py:285).  Instead of an NCCL all-reduce followed
// by the norm kernel, ONE kernel per rank does the exchange and the first half of the optimiser step over NVLink peer
// memory: it publishes "my gradient of step s is complete" to every peer, waits for all peers, reads every rank's gradient
// (+ the 4 loss sums) straight from that rank's HBM over NVLink, adds them in rank order 0..G-1 -- so every rank obtains
// the bit-identical sum -- writes the reduced gradient locally and emits the partial sums of squares of the global norm.
// clip_adam_kernel then finishes the step on every replica.  861 KB per rank at config S: latency-bound (one NVLink round
// trip), which is why it is one fused launch rather than a ring.
//
// Exchange buffer per rank (cudaMalloc'ed here because CUDA IPC handles need a base allocation; everything else in libdppo
// runs on caller-owned memory): [slot 0 | slot 1 | flags], slot = gradient (n floats) + 4 loss sums, padded.  Steps alternate
// slots; a slot is re-written two steps later, which the flag protocol orders after every peer's read (a peer signals
// exchange s+1 only after its exchange-s kernel, stream order).  flags[q] = last exchange rank q has published.
// Exchanges are numbered by their own monotonic sequence (1, 2, ...; the caller passes it), NOT by the Adam step: restoring an
// optimiser checkpoint may move the Adam step backwards, the sequence never does.  The wait is bounded: a peer that does not
// publish within ~10 s (crashed or desynchronised rank) makes the kernel give up and raise a flag in host-mapped memory that
// dppo_dp_status() reports, instead of hanging the GPU.
#include <cuda_runtime.h>

#include "common.cuh"
#include "optim.cuh"

struct dppo_dp {
    int world, rank;
    int64_t n, slot_floats;
    float* local;                       // this rank's exchange buffer
    float* peer[DPPO_MAX_RANKS];        // every rank's buffer as mapped in this process (peer[rank] == local)
    int opened[DPPO_MAX_RANKS];
    cudaIpcMemHandle_t handle;
    int* status_host;                   // host-mapped: 0 ok, else 1 + rank the kernel timed out waiting for
    int* status_dev;
};

namespace {

struct PeerPtrs { const float* slot[DPPO_MAX_RANKS]; unsigned long long* flags[DPPO_MAX_RANKS]; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int DP_THREADS = 256;
constexpr long long DP_WAIT_CYCLES = 20000000000ll;      // ~10 s at 1.9 GHz

// n4: float4 elements of the gradient; the 4 loss sums follow as one more float4 (excluded from the norm)
__global__ void __launch_bounds__(DP_THREADS)
dp_allreduce_sumsq_kernel(PeerPtrs pp, int world, int rank, unsigned long long step, const unsigned long long* __restrict__ step_dev,
                          int64_t n4, float* __restrict__ grads_out, float* __restrict__ losses_out, double* __restrict__ partials,
                          int* __restrict__ status)
{
    __shared__ double red[DP_THREADS / 32];
    if (step_dev != nullptr) step = *step_dev;      // CUDA-graph replay: the sequence number is device-resident
    // publish: the gradient of this step was written by earlier kernels of this stream, i.e. it is complete
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(pp.flags[threadIdx.x] + rank, step);
    }
    // wait until every rank has published this step (flags live in this rank's own buffer)
    if (threadIdx.x < world) {
        const unsigned long long* f = pp.flags[rank] + threadIdx.x;
        const long long t0 = clock64();
        while (ld_acquire_sys(f) < step) {
            __nanosleep(20);
            if (clock64() - t0 > DP_WAIT_CYCLES) {             // peer never published: report instead of hanging the GPU
                if (blockIdx.x == 0) *status = 1 + (int)threadIdx.x;
                break;
            }
        }
    }
    __syncthreads();
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n4; i += (int64_t)gridDim.x * blockDim.x) {
        // every rank's element is requested before the first one is used: ONE NVLink round trip per thread instead of `world`
        // dependent ones (the loop over a run-time rank count serialised load -> add -> load: 8 x ~2.5 us at 8 GPUs)
        float4 v[DPPO_MAX_RANKS];
#pragma unroll
        for (int q = 0; q < DPPO_MAX_RANKS; ++q)
            if (q < world) v[q] = __ldcv(reinterpret_cast<const float4*>(pp.slot[q]) + i);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < DPPO_MAX_RANKS; ++q)                             // fixed rank order: identical sums on every rank
            if (q < world) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        if (i < n4) {
            reinterpret_cast<float4*>(grads_out)[i] = acc;
            s += (double)acc.x * acc.x + (double)acc.y * acc.y + (double)acc.z * acc.z + (double)acc.w * acc.w;
        } else if (losses_out) {
            *reinterpret_cast<float4*>(losses_out) = acc;
        }
    }
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < DP_THREADS / 32 ? red[threadIdx.x] : 0.0;
        s = warp_sum_d(s);
        if (threadIdx.x == 0) partials[blockIdx.x] = s;
    }
}

int dp_blocks(int64_t n4)
{
    int64_t b = (n4 + 1 + DP_THREADS - 1) / DP_THREADS;
    if (b > 256) b = 256;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" int dppo_dp_create(dppo_ctx* ctx, int world, int rank, int64_t n_floats, dppo_dp** out)
{
    if (!ctx || !out) return 1;
    *out = nullptr;
    if (world < 1 || world > DPPO_MAX_RANKS || rank < 0 || rank >= world) DPPO_FAIL(ctx, "dp_create: bad world/rank %d/%d", world, rank);
    if (n_floats <= 0 || n_floats % 4 != 0) DPPO_FAIL(ctx, "dp_create: gradient length must be a positive multiple of 4 floats");
    dppo_dp* dp = new dppo_dp();
    dp->world = world; dp->rank = rank; dp->n = n_floats;
    dp->slot_floats = align_up(n_floats + 4, 64);
    const size_t bytes = (size_t)(2 * dp->slot_floats) * 4 + DPPO_MAX_RANKS * sizeof(unsigned long long);
    for (int q = 0; q < DPPO_MAX_RANKS; ++q) { dp->peer[q] = nullptr; dp->opened[q] = 0; }
    dp->status_host = nullptr; dp->status_dev = nullptr;
    cudaError_t e = cudaMalloc((void**)&dp->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(dp->local, 0, bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&dp->handle, dp->local);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&dp->status_host, sizeof(int), cudaHostAllocMapped);
    if (e == cudaSuccess) { *dp->status_host = 0; e = cudaHostGetDevicePointer((void**)&dp->status_dev, dp->status_host, 0); }
    if (e != cudaSuccess) {
        if (dp->status_host) cudaFreeHost(dp->status_host);
        if (dp->local) cudaFree(dp->local);
        delete dp;
        DPPO_FAIL(ctx, "dp_create: %s", cudaGetErrorString(e));
    }
    dp->peer[rank] = dp->local;
    *out = dp;
    return 0;
}

extern "C" int dppo_dp_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

extern "C" int dppo_dp_handle(dppo_dp* dp, void* handle_out)
{
    if (!dp || !handle_out) return 1;
    memcpy(handle_out, &dp->handle, sizeof(cudaIpcMemHandle_t));
    return 0;
}

// all_handles: world consecutive cudaIpcMemHandle_t, index = rank (as all-gathered by the caller)
extern "C" int dppo_dp_connect(dppo_ctx* ctx, dppo_dp* dp, const void* all_handles)
{
    if (!ctx || !dp || !all_handles) return 1;
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int q = 0; q < dp->world; ++q) {
        if (q == dp->rank) continue;
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) DPPO_FAIL(ctx, "dp_connect: cudaIpcOpenMemHandle(rank %d): %s", q, cudaGetErrorString(e));
        dp->peer[q] = static_cast<float*>(p);
        dp->opened[q] = 1;
    }
    return 0;
}

extern "C" int dppo_dp_destroy(dppo_dp* dp)
{
    if (!dp) return 0;
    for (int q = 0; q < dp->world; ++q)
        if (dp->opened[q]) cudaIpcCloseMemHandle(dp->peer[q]);
    if (dp->local) cudaFree(dp->local);
    if (dp->status_host) cudaFreeHost(dp->status_host);
    delete dp;
    return 0;
}

// 0: every exchange so far found its peers; 1 + q: an exchange kernel gave up waiting for rank q (no device synchronisation:
// the kernels write the flag to host-mapped memory)
extern "C" int dppo_dp_status(dppo_dp* dp) { return dp && dp->status_host ? *(volatile int*)dp->status_host : -1; }

// Device pointer the LOCAL gradient (n floats) and, right after it, the 4 local loss sums of exchange number `seq` must be
// written to (pass it as `grads` / `losses` to dppo_mlp_grad_minibatch).
extern "C" float* dppo_dp_slot(dppo_dp* dp, int64_t seq) { return dp ? dp->local + (seq & 1) * dp->slot_floats : nullptr; }

extern "C" int64_t dppo_dp_workspace_bytes(int64_t n) { return (int64_t)dp_blocks(n / 4) * (int64_t)sizeof(double); }

// Fused exchange + optimiser step: grads_out (n floats, local) receives the rank-ordered sum of every rank's slot, losses_out
// (4 floats, optional) the summed loss sums; then clip_grad_norm_ + Adam run on this replica exactly as dppo_clip_adam_step.
extern "C" int dppo_dp_allreduce_clip_adam(dppo_ctx* ctx, dppo_dp* dp, int64_t seq, float* params, float* grads_out, float* exp_avg,
                                           float* exp_avg_sq, const dppo_hyper* h, float* losses_out, float* grad_norm_out,
                                           void* ws, int64_t ws_bytes, void* stream)
{
    if (!ctx) return 1;
    if (!dp || !params || !grads_out || !exp_avg || !exp_avg_sq || !h || !ws) DPPO_FAIL(ctx, "dp_allreduce_clip_adam: null argument");
    if (h->step < 1 || seq < 1) DPPO_FAIL(ctx, "dp_allreduce_clip_adam: step and seq must be >= 1");
    for (int q = 0; q < dp->world; ++q)
        if (!dp->peer[q]) DPPO_FAIL(ctx, "dp_allreduce_clip_adam: rank %d is not connected (dppo_dp_connect)", q);
    const int64_t n4 = dp->n / 4;
    const int nb = dp_blocks(n4);
    if (ws_bytes < (int64_t)nb * (int64_t)sizeof(double)) DPPO_FAIL(ctx, "dp_allreduce_clip_adam: workspace too small");
    if ((reinterpret_cast<uintptr_t>(grads_out) & 15u) || (losses_out && (reinterpret_cast<uintptr_t>(losses_out) & 15u)))
        DPPO_FAIL(ctx, "dp_allreduce_clip_adam: grads_out / losses_out must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    PeerPtrs pp;
    const int64_t slot_off = (seq & 1) * dp->slot_floats;
    for (int q = 0; q < DPPO_MAX_RANKS; ++q) {
        pp.slot[q] = q < dp->world ? dp->peer[q] + slot_off : nullptr;
        pp.flags[q] = q < dp->world ? reinterpret_cast<unsigned long long*>(dp->peer[q] + 2 * dp->slot_floats) : nullptr;
    }
    double* partials = (double*)ws;
    // h->step_consts (optional, device): {2 floats for the Adam kernel, then the 64-bit exchange sequence number of this launch}
    const unsigned long long* step_dev = h->step_consts ? reinterpret_cast<const unsigned long long*>(h->step_consts + 2) : nullptr;
    dp_allreduce_sumsq_kernel<<<nb, DP_THREADS, 0, st>>>(pp, dp->world, dp->rank, (unsigned long long)seq, step_dev, n4, grads_out,
                                                         losses_out, partials, dp->status_dev);
    DPPO_CHECK_LAUNCH(ctx, "dp_allreduce_sumsq_kernel");
    return launch_clip_adam(ctx, params, grads_out, exp_avg, exp_avg_sq, dp->n, partials, nb, h, grad_norm_out, st);
}

// A rank that owns no row of a global minibatch contributes zeros
extern "C" int dppo_dp_zero_slot(dppo_ctx* ctx, dppo_dp* dp, int64_t seq, void* stream)
{
    if (!ctx || !dp) return 1;
    if (cudaMemsetAsync(dppo_dp_slot(dp, seq), 0, (size_t)dp->slot_floats * 4, (cudaStream_t)stream) != cudaSuccess)
        DPPO_FAIL(ctx, "dp_zero_slot: cudaMemsetAsync failed");
    return 0;
}
