// Shared declarations for libdppo's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "dppo.h"

struct dppo_ctx {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    char err[512];
    int use_tensor_cores;                 // != 0: 3xTF32 tcgen05 GEMMs where the shape allows (default); 0: FP32 FFMA GEMMs only
    int gae_variant;                      // 0: auto (pipelined TMA kernel for T >= 128), 1: register-staged kernel, 2: single-barrier TMA kernel
    int gae_inputs_settled;               // 1: caller guarantees the GAE inputs are not written by the kernel just before the GAE launch
    long long launch_count;               // kernels launched through this context (bench.py's gpu_launches)
    int row_sweep;                        // L2 reuse between consecutive launches of the layer chain (gemm_tc3.cu dppo_tc3_gemm), default 31.
                                          // bit 0: consecutive GEMM launches alternate their row-sweep direction, bit 1: the head kernels
                                          // sweep against the GEMM before them, bit 2: the head kernel's d3 stores stay cacheable, bit 3:
                                          // GEMM / head-kernel inputs are read with the L2 evict-first hint, bit 4: so are the operands of
                                          // the weight-gradient launch.  Bits 2-4 do not change any result; bits 0 / 1 re-order the fp32
                                          // partial sums of the bias gradients / the head kernel's accumulators (forward outputs unchanged)
    int tc_prefetch;                      // software L2 prefetch, bit mask: 1 forward GEMMs / 2 dgrad GEMMs (next tile's activations), 4 weight
                                          // gradient (operand chunks 8 ahead); default 0: every one of them measured slower in the graph replay (gemm_tc3.cu)
    int tc_debug;                         // timing-experiment switches; only honoured by builds with -DDPPO_TIMING_SWITCHES (see DPPO_DBG)
    const unsigned long long* draw_base;  // optional device counter added to every sampling draw counter (CUDA-graph replay of rollouts)
    const int* rows_dev;                  // optional DEVICE row count: while set, the forward kernels (gather, GEMMs, head evaluation) process
                                          // min(rows, *rows_dev) rows -- shapes decided on the device, launches of fixed size, no host sync
    void* tm_cache;                       // tensor-map cache owned by gae.cu
    void (*tm_cache_free)(void*);
};

extern char g_dppo_create_error[512];

#define DPPO_FAIL(ctx, ...)                                   \
    do {                                                      \
        if (ctx) snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__); \
        return 1;                                             \
    } while (0)

#define DPPO_CHECK_LAUNCH(ctx, what)                                                        \
    do {                                                                                    \
        cudaError_t e_ = cudaGetLastError();                                                \
        if (e_ != cudaSuccess) DPPO_FAIL(ctx, "%s: %s", what, cudaGetErrorString(e_));      \
        ++(ctx)->launch_count;                                                              \
    } while (0)

// Work-skipping / A-B switches of the tensor-core kernels (skip weight copies, single-pass TF32, no epilogue, ...) give WRONG
// results and exist for timing experiments only.  They are compiled out of the release library: DPPO_DBG() is a constant
// false unless the library is built with `make EXTRA=-DDPPO_TIMING_SWITCHES`, and dppo_set_option("tc_debug") then fails.
#ifdef DPPO_TIMING_SWITCHES
#define DPPO_DBG(mask, bit) (((mask) & (bit)) != 0)
#else
#define DPPO_DBG(mask, bit) false
#endif

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

// ---- programmatic dependent launch along the kernels of an optimiser step (opt-in: timing builds, tc_debug bit 512) ---------------
// MEASURED NEGATIVE on B200 for this chain (graph replay, config S: 605-615 us per optimiser step with the attribute,
// 580-600 us without), so the attribute is off by default and the instructions below are no-ops; kept for A/B.
// With the bit set every kernel of the update loop is launched with programmatic stream serialisation and starts with DPPO_PDL_ENTER():
// it releases its own dependents at once and then waits for its predecessor grid to complete (and its writes to become
// visible) before touching global memory.  The work itself stays fully ordered; what overlaps is the launch latency and the
// per-CTA set-up (barrier init, TMEM allocation, tensor-map fetch) of kernel i+1 with the tail of kernel i: CTAs of the
// next grid become resident as soon as an SM is free.  A kernel launched without the attribute sees both instructions as
// no-ops.  Every kernel launched with the attribute MUST execute the wait (completion of a grid then implies completion of
// all its predecessors).
#define DPPO_PDL_ENTER()                                                    \
    do {                                                                    \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     \
        asm volatile("griddepcontrol.wait;" ::: "memory");                  \
    } while (0)

// on: launch with programmatic stream serialisation (the kernel must start with DPPO_PDL_ENTER())
template <typename... KArgs, typename... Args>
static inline cudaError_t dppo_launch_pdl_if(bool on, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = on ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
static inline cudaError_t dppo_launch_pdl(dppo_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                          Args&&... args)
{
    return dppo_launch_pdl_if(DPPO_DBG(ctx->tc_debug, 512), kern, grid, block, smem, st, static_cast<Args&&>(args)...);
}

// L2 eviction-priority hint for data that is dead after this read (row_sweep bit 3 / 4): at the end of a launch the L2 should hold
// the launch's OUTPUT, which the next launch starts with, not a mix of output and consumed input.
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_l2_hint(const float4* p, uint64_t policy)
{
    float4 v;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- internal launchers shared between translation units ----------------------------------
// C[M,N] = epi(A[M,K] * op(B)); see gemm_simt.cu
enum { DPPO_EPI_BIAS = 0, DPPO_EPI_BIAS_TANH = 1, DPPO_EPI_TANH_BWD = 2, DPPO_EPI_NONE = 3 };
// NT: B is [N,K] row-major (a torch Linear weight used in the forward direction)
int dppo_gemm_nt(dppo_ctx* ctx, int epi, const float* A, int lda, const int32_t* a_rows, const float* B, int ldb,
                 const float* bias, float* C, int ldc, int64_t M, int N, int K, cudaStream_t st);
// NN: B is [K,N] row-major (a Linear weight used in the backward/dgrad direction);
// C = (A*B) .* (1 - Hact^2); colsum (optional) receives per-row-tile column sums of C: [tiles_m, N]
int dppo_gemm_nn_tanh_bwd(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, const float* Hact, int ldh,
                          float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st);
// plain C = A*B with B [K,N] row-major (gradient w.r.t. the input of a Linear layer that is not followed by tanh')
int dppo_gemm_nn(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                 cudaStream_t st);
int dppo_gemm_row_tiles(int64_t M, int N);     // number of row tiles the NN kernel uses (colsum partial count)
// TN split-K weight gradient: partials[s][n1][n2] = sum over the rows of split s of A[m,n1]*B[m,n2]
int dppo_wgrad_splits(dppo_ctx* ctx, int64_t M, int N1, int N2);
int dppo_wgrad(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, const int32_t* b_rows,
               float* partials, int splits, int64_t M, int N1, int N2, cudaStream_t st);
