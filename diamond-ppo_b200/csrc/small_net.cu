// Persistent update kernel for the default 64-wide actor-critic MLP (BASELINE.json north_star item 4; diamond/ppo.py:258-285).
//
// At the reference's default network size (hidden 64: 13 k parameters, 128..512-row minibatches) one optimiser step is ~13 MFLOP:
// under a microsecond of math on a B200, but ten kernel launches (gather, three GEMMs, heads, two dgrads, wgrads, reduction, clip
// + Adam) cost ~57 us even when replayed as a CUDA graph.  Here ALL optimiser steps of a learn() -- every epoch, every minibatch --
// run inside ONE launch of one thread-block cluster of 8 CTAs:
//   * every CTA keeps a full copy of the parameters in shared memory (52 KB) and takes 1/8 of the minibatch rows, 16 at a time:
//     gather -> forward -> heads + PPO loss -> backward -> weight gradients, all on shared-memory tiles with FP32 FFMA (activations
//     are kept k-major so that a warp's 16 rows read consecutive words; every gradient element is owned by exactly one thread, so
//     the CTA's partial gradient accumulates in shared memory without atomics, in a fixed order);
//   * reduce-scatter over distributed shared memory: CTA c sums slice c of the eight partial gradients (fixed rank order) and the
//     slice's sum of squares; after a cluster barrier every CTA adds the eight partial norms (same order -> same clip factor);
//   * CTA c applies clip_grad_norm_ + Adam to its slice (it alone holds that slice of exp_avg / exp_avg_sq) and writes the new
//     parameter values into all eight parameter copies (all-gather over DSMEM); one more cluster barrier ends the step.
// Three cluster barriers per optimiser step, no global-memory traffic except the gathered rollout rows and 16 bytes of losses.
// The arithmetic per element is the one of heads.cu / optim.cu (torch/nn/utils/clip_grad.py:165-182, torch/optim/adam.py:531-547).
#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

#ifdef DPPO_SMALL_PROFILE
long long* g_small_prof = nullptr;
extern "C" long long* dppo_small_prof(void) { return g_small_prof; }
#endif

namespace {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;
// RT = rows per tile (template parameter): 32 (lane <-> row) or 16 for the smallest minibatches (two column groups per warp), so that
// a CTA holding only 16 rows of the minibatch still uses every lane.
template <int RT> struct RowStride { static constexpr int v = RT + 4; };   // row stride of the k-major activation tiles (floats):
                                    // 16-byte aligned rows, and consecutive rows start 4 banks apart, so a warp reading 32 different
                                    // rows with float4 loads is conflict-free
constexpr int HN = 64;              // hidden width this kernel is specialised for
constexpr int AMAX = 8;             // actions / action dims
constexpr int ROWS_MAX = 128;       // minibatch rows per CTA

struct SmallArgs {
    dppo_mlp_layout L;
    int D, A;
    float* P; float* G; float* M; float* V;                 // flat global buffers (L.total floats)
    const float* obs; const void* actions; const float* old_logp; const float* adv; const float* ret;
    const double* adv_stats; long long adv_count; int advantage_norm;
    const int32_t* idx;                                     // [steps][rows]
    int rows, steps;
    const float* step_consts;                               // [steps][2]: sqrt(1 - beta2^t), -lr / (1 - beta1^t)
    float clip, vw, beta, inv_m, max_norm, w1, beta2, w2, eps;
    float* losses;                                          // [steps][4]
    float* grad_norm_out;                                   // optional: pre-clip norm of the last step
    long long* prof;                                        // DPPO_SMALL_PROFILE builds: per-phase cycle counts of CTA 0
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// out[n][r] = tanh(b[n] + sum_k in[k][r] * W[n][k]) for the 32 rows of a tile; lane = row, warp = group of N/8 output columns
// (the weight reads are warp-wide broadcasts, the activation reads conflict-free)
template <int N, int RT>
__device__ __forceinline__ void fwd_layer(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ b,
                                          float* __restrict__ out)
{
    constexpr int RS = RowStride<RT>::v;
    constexpr int NC = N * RT / SK_THREADS;           // columns per thread: thread = (row r, column group)
    const int r = threadIdx.x % RT, c0 = (threadIdx.x / RT) * NC;
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = b[c0 + j];
    if ((K & 3) == 0) {
#pragma unroll 2
        for (int k = 0; k < K; k += 4) {
            const float a0 = in[k * RS + r], a1 = in[(k + 1) * RS + r], a2 = in[(k + 2) * RS + r], a3 = in[(k + 3) * RS + r];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 w = ld4(W + (c0 + j) * K + k);
                acc[j] = fmaf(a0, w.x, acc[j]); acc[j] = fmaf(a1, w.y, acc[j]); acc[j] = fmaf(a2, w.z, acc[j]); acc[j] = fmaf(a3, w.w, acc[j]);
            }
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float a = in[k * RS + r];
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, W[(c0 + j) * K + k], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) out[(c0 + j) * RS + r] = tanhf(acc[j]);
}

// din[k][r] = (sum_n dout[n][r] * W[n][k]) * (1 - hin[k][r]^2); lane = row, warp = 8 consecutive k; W row-major [N][64]
template <int N, int RT>
__device__ __forceinline__ void dgrad_layer(const float* __restrict__ dout, const float* __restrict__ W, const float* __restrict__ hin,
                                            float* __restrict__ din)
{
    constexpr int RS = RowStride<RT>::v;
    constexpr int KC = HN * RT / SK_THREADS;          // outputs per thread: 8 (RT = 32) or 4 (RT = 16)
    const int r = threadIdx.x % RT, k0 = (threadIdx.x / RT) * KC;
    float acc[KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) acc[q] = 0.f;
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float d = dout[n * RS + r];
#pragma unroll
        for (int q = 0; q < KC; q += 4) {
            const float4 w = ld4(W + n * HN + k0 + q);
            acc[q] = fmaf(d, w.x, acc[q]); acc[q + 1] = fmaf(d, w.y, acc[q + 1]); acc[q + 2] = fmaf(d, w.z, acc[q + 2]); acc[q + 3] = fmaf(d, w.w, acc[q + 3]);
        }
    }
#pragma unroll
    for (int q = 0; q < KC; ++q) {
        const float h = hin[(k0 + q) * RS + r];
        din[(k0 + q) * RS + r] = acc[q] * (1.0f - h * h);
    }
}

template <int RT>
__device__ __forceinline__ float dotR(const float* __restrict__ a, const float* __restrict__ b)
{
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int q = 0; q < RT; q += 8) {
        const float4 x = ld4(a + q), y = ld4(b + q), x2 = ld4(a + q + 4), y2 = ld4(b + q + 4);
        s0 = fmaf(x.x, y.x, s0); s0 = fmaf(x.y, y.y, s0); s0 = fmaf(x.z, y.z, s0); s0 = fmaf(x.w, y.w, s0);
        s1 = fmaf(x2.x, y2.x, s1); s1 = fmaf(x2.y, y2.y, s1); s1 = fmaf(x2.z, y2.z, s1); s1 = fmaf(x2.w, y2.w, s1);
    }
    return s0 + s1;
}
template <int RT>
__device__ __forceinline__ float sumR(const float* __restrict__ a)
{
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RT; q += 4) { const float4 x = ld4(a + q); s += (x.x + x.y) + (x.z + x.w); }
    return s;
}

// g[n][k] += sum_r dout[n][r] * hin[k][r] for a [N][64] weight: thread = (k = tid % 64, quarter of the n range); the thread keeps
// its activation row hin[k][0..31] in registers, the gradient rows are warp-wide broadcasts, g is written with consecutive k
template <int N, int RT>
__device__ __forceinline__ void wgrad_layer(const float* __restrict__ dout, const float* __restrict__ hin, float* __restrict__ g)
{
    constexpr int RS = RowStride<RT>::v;
    const int k = threadIdx.x & 63, n0 = (threadIdx.x >> 6) * (N / 4);
    float4 h[RT / 4];
#pragma unroll
    for (int q = 0; q < RT / 4; ++q) h[q] = ld4(hin + k * RS + 4 * q);
#pragma unroll 2
    for (int n = n0; n < n0 + N / 4; ++n) {
        const float* d = dout + n * RS;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int q = 0; q < RT / 4; q += 2) {
            const float4 x = ld4(d + 4 * q), y = ld4(d + 4 * q + 4);
            s0 = fmaf(x.x, h[q].x, s0); s0 = fmaf(x.y, h[q].y, s0); s0 = fmaf(x.z, h[q].z, s0); s0 = fmaf(x.w, h[q].w, s0);
            s1 = fmaf(y.x, h[q + 1].x, s1); s1 = fmaf(y.y, h[q + 1].y, s1); s1 = fmaf(y.z, h[q + 1].z, s1); s1 = fmaf(y.w, h[q + 1].w, s1);
        }
        g[n * HN + k] += s0 + s1;
    }
}

__device__ __forceinline__ float group8_sum(float v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 4, 8); v += __shfl_xor_sync(0xffffffffu, v, 2, 8); v += __shfl_xor_sync(0xffffffffu, v, 1, 8);
    return v;
}
__device__ __forceinline__ float group8_max(float v)
{
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4, 8)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2, 8));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1, 8));
    return v;
}

#ifdef DPPO_SMALL_PROFILE
#define SK_MARK(i) do { if (prof_on) { __syncthreads(); if (tid == 0) { const long long now = clock64(); a.prof[i] += now - prof_t; prof_t = now; } } } while (0)
#else
#define SK_MARK(i) do { } while (0)
#endif

template <bool CONT, int RT, int SK_CL>
__global__ void __launch_bounds__(SK_THREADS, 1)
small_update_kernel(const SmallArgs a)
{
    constexpr int RS = RowStride<RT>::v;
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const dppo_mlp_layout& L = a.L;
    const int total = (int)L.total, D = a.D, A = a.A;
    const int SL = ((total + SK_CL - 1) / SK_CL + 3) & ~3;              // slice of the flat buffers owned by one CTA
    const int s0 = rank * SL, s1 = min(total, s0 + SL);

    float* sP = smem;                          // [total] parameters (identical in all CTAs)
    float* sG = sP + total;                    // [total] this CTA's gradient partial
    float* sGs = sG + total;                   // [SL] reduced gradient slice
    float* sMs = sGs + SL;                     // [SL] exp_avg slice
    float* sVs = sMs + SL;                     // [SL] exp_avg_sq slice
    float* sX = sVs + SL;                      // [D][RS]
    float* sH1 = sX + D * RS;                  // [64][RS]
    float* sH2 = sH1 + HN * RS;
    float* sH3 = sH2 + HN * RS;                // [128][RS]
    float* sD3 = sH3 + 2 * HN * RS;
    float* sD2 = sD3 + 2 * HN * RS;
    float* sD1 = sD2 + HN * RS;
    float* sZ = sD1 + HN * RS;                 // [RT][AMAX] head outputs
    float* sVal = sZ + RT * AMAX;              // [RT] values
    float* sDz = sVal + RT;                    // [AMAX + 1][RS]: d(loss)/dz per action (rows 0..A-1), d(loss)/dv (row AMAX)
    float* sDls = sDz + (AMAX + 1) * RS;       // [AMAX][RS] per-row d(loss)/d(log_std) (continuous)
    float* sRed = sDls + AMAX * RS;            // [8]: [0..2] loss sums, [4..5] norm partial (double)
    __shared__ int s_src[ROWS_MAX], s_acti[ROWS_MAX];
    __shared__ float s_oldlp[ROWS_MAX], s_adv[ROWS_MAX], s_ret[ROWS_MAX], s_actf[ROWS_MAX * AMAX];
    __shared__ float s_coef;
    __shared__ float s_loss[3][SK_WARPS];
    __shared__ double s_wsum[SK_WARPS];

    for (int i = tid; i < total; i += SK_THREADS) { sP[i] = a.P[i]; sG[i] = 0.f; }
    for (int i = s0 + tid; i < s1; i += SK_THREADS) { sMs[i - s0] = a.M[i]; sVs[i - s0] = a.V[i]; }
    float adv_mean = 0.f, adv_denom = 1.f;
    if (a.advantage_norm) {
        const double mu = a.adv_stats[0] / (double)a.adv_count;
        double var = (a.adv_stats[1] - a.adv_stats[0] * mu) / (double)(a.adv_count - 1);
        var = var > 0.0 ? var : 0.0;
        adv_mean = (float)mu;
        adv_denom = (float)sqrt(var) + 1e-6f;                         // ppo.py:243
    }
    const int rows_per = (a.rows + SK_CL - 1) / SK_CL;
    const int row_lo = min(a.rows, rank * rows_per), row_hi = min(a.rows, row_lo + rows_per);
    const int my_rows = row_hi - row_lo;
    int src_next = -1;                                                   // my row's index for the NEXT step (loaded one step ahead)
    cluster.sync();

#ifdef DPPO_SMALL_PROFILE
    const bool prof_on = a.prof != nullptr && rank == 0;
    long long prof_t = clock64();
#endif
    for (int step = 0; step < a.steps; ++step) {
        SK_MARK(15);
        const int32_t* idx = a.idx + (int64_t)step * a.rows + row_lo;
        // ---- per-row scalars of all my rows of this step (ppo.py:265-272) ----
        // index -> data is a chain of two global-memory latencies (~4 us per step when exposed); every step's lists are known up
        // front, so thread m keeps the pipeline two steps deep in registers: its index for step + 1 was loaded a step ago, the data
        // loads for step + 1 are issued now and land in shared memory at the end of this step.
        if (step == 0) {
            if (tid < my_rows) {
                const int src = idx[tid];
                const bool live = src >= 0;
                s_src[tid] = src;
                s_oldlp[tid] = live ? a.old_logp[src] : 0.f;
                s_adv[tid] = live ? (a.adv[src] - adv_mean) / adv_denom : 0.f;
                s_ret[tid] = live ? a.ret[src] : 0.f;
                if (!CONT) s_acti[tid] = live ? reinterpret_cast<const int32_t*>(a.actions)[src] : 0;
                else
                    for (int j = 0; j < A; ++j) s_actf[tid * AMAX + j] = live ? reinterpret_cast<const float*>(a.actions)[(int64_t)src * A + j] : 0.f;
                src_next = a.steps > 1 ? idx[a.rows + tid] : -1;
            }
        }
        float n_oldlp = 0.f, n_adv = 0.f, n_ret = 0.f, n_actf[AMAX];
        int n_acti = 0, n_src = -1, src_next2 = -1;
        if (tid < my_rows && step + 1 < a.steps) {
            n_src = src_next;
            const bool live = n_src >= 0;
            n_oldlp = live ? a.old_logp[n_src] : 0.f;
            n_adv = live ? a.adv[n_src] : adv_mean;
            n_ret = live ? a.ret[n_src] : 0.f;
            if (!CONT) n_acti = live ? reinterpret_cast<const int32_t*>(a.actions)[n_src] : 0;
            else {
#pragma unroll
                for (int j = 0; j < AMAX; ++j) n_actf[j] = (live && j < A) ? reinterpret_cast<const float*>(a.actions)[(int64_t)n_src * A + j] : 0.f;
            }
            if (step + 2 < a.steps) src_next2 = idx[2 * (int64_t)a.rows + tid];
        }
        float l_pol = 0.f, l_val = 0.f, l_ent = 0.f;                      // threads with (tid & 7) == 0: sums over their rows
        __syncthreads();
        SK_MARK(0);

        for (int t0 = 0; t0 < my_rows; t0 += RT) {
            // ---- gather the observations of the tile (ppo.py:261) ----
            for (int i = tid; i < D * RT; i += SK_THREADS) {
                const int r = i / D, k = i - r * D;                      // consecutive threads read consecutive floats of a row
                const int src = t0 + r < my_rows ? s_src[t0 + r] : -1;
                sX[k * RS + r] = src >= 0 ? a.obs[(int64_t)src * D + k] : 0.f;
            }
            __syncthreads();
            SK_MARK(1);
            // ---- forward (ppo.py:91-96) ----
            fwd_layer<HN, RT>(sX, D, sP + L.w1, sP + L.b1, sH1);
            __syncthreads();
            SK_MARK(2);
            fwd_layer<HN, RT>(sH1, HN, sP + L.w2, sP + L.b2, sH2);
            __syncthreads();
            SK_MARK(3);
            fwd_layer<2 * HN, RT>(sH2, HN, sP + L.w3, sP + L.b3, sH3);
            __syncthreads();
            SK_MARK(4);
            {   // output heads: thread = (row r, output o): o < A actor, o == A critic
                const int r = tid % RT;
                for (int o = tid / RT; o <= A; o += SK_THREADS / RT) {
                    const float* w = o < A ? sP + L.wa + o * HN : sP + L.wc;
                    const float* h = o < A ? sH3 : sH3 + HN * RS;
                    float s = o < A ? sP[L.ba + o] : sP[L.bc];
#pragma unroll 4
                    for (int k = 0; k < HN; k += 4) {
                        const float4 w4 = ld4(w + k);
                        s = fmaf(h[k * RS + r], w4.x, s); s = fmaf(h[(k + 1) * RS + r], w4.y, s);
                        s = fmaf(h[(k + 2) * RS + r], w4.z, s); s = fmaf(h[(k + 3) * RS + r], w4.w, s);
                    }
                    if (o < A) sZ[r * AMAX + o] = s; else sVal[r] = s;
                }
            }
            __syncthreads();
            SK_MARK(5);
            // ---- distribution, loss terms, d(loss)/d(head outputs) (ppo.py:264-280): 8 lanes per row, lane j <-> action j ----
            {
                const int r = tid >> 3, j = tid & 7, m = t0 + r;           // (RT = 16: the upper half of the block idles here)
                const bool in_tile = r < RT;
                const bool live = in_tile && m < my_rows && s_src[m] >= 0;
                const bool on = j < A;
                const float z = (on && in_tile) ? sZ[r * AMAX + j] : 0.f;
                float new_lp, entropy, dz = 0.f, dls = 0.f;
                float p = 0.f, lsm = 0.f, diff = 0.f, var = 1.f;
                const int act = (!CONT && live) ? s_acti[m] : 0;
                if (!CONT) {
                    // torch/distributions/categorical.py:78 (logits - logsumexp), :151-163 (log_prob, entropy)
                    const float mx = group8_max(on ? z : -CUDART_INF_F);
                    const float s = group8_sum(on ? expf(z - mx) : 0.f);
                    const float lse = mx + logf(s);
                    lsm = on ? z - lse : 0.f;
                    p = on ? expf(lsm) : 0.f;
                    entropy = -group8_sum(p * lsm);
                    new_lp = __shfl_sync(0xffffffffu, lsm, act, 8);
                } else {
                    // torch/distributions/normal.py:87-102, :114-115, summed over dims (continuous_ppo.py:40-47)
                    const float sigma = on ? expf(sP[L.log_std + j]) : 1.f;
                    const float log_scale = logf(sigma);
                    var = sigma * sigma;
                    diff = (on && live ? s_actf[m * AMAX + j] : 0.f) - z;
                    new_lp = group8_sum(on ? (-(diff * diff) / (2.0f * var) - log_scale - 0.91893853320467274178f) : 0.f);
                    entropy = group8_sum(on ? (0.5f + 0.91893853320467274178f + log_scale) : 0.f);
                }
                // clipped surrogate (ppo.py:266-270) and the gradient autograd assigns
                const float adv = live ? s_adv[m] : 0.f;
                const float ratio = expf(new_lp - (live ? s_oldlp[m] : 0.f));
                const float lo = 1.0f - a.clip, hi = 1.0f + a.clip;
                const float u1 = -adv * ratio, u2 = -adv * fminf(fmaxf(ratio, lo), hi);
                const bool in_range = ratio >= lo && ratio <= hi;
                const float wsel = in_range ? 1.0f : (u1 > u2 ? 1.0f : (u1 == u2 ? 0.5f : 0.0f));
                const float dlogp = (-adv * wsel * a.inv_m) * ratio;
                const float verr = (in_tile ? sVal[r] : 0.f) - (live ? s_ret[m] : 0.f);
                if (live && on) {
                    if (!CONT) dz = dlogp * ((j == act ? 1.0f : 0.0f) - p) + (a.beta * a.inv_m) * p * (lsm + entropy);
                    else { dz = dlogp * diff / var; dls = dlogp * (diff * diff / var - 1.0f) - a.beta * a.inv_m; }
                }
                if (in_tile) {
                    sDz[j * RS + r] = dz;
                    if (CONT) sDls[j * RS + r] = dls;
                }
                if (j == 0 && in_tile) {
                    sDz[AMAX * RS + r] = live ? a.vw * verr * a.inv_m : 0.f;
                    if (live) { l_pol += fmaxf(u1, u2); l_val += verr * verr; l_ent += entropy; }
                }
            }
            __syncthreads();
            SK_MARK(6);
            // ---- backward into the first head layers: d3[n][r]; thread = (row r, group of 128 * RT / 256 columns) ----
            {
                constexpr int NC3 = 2 * HN * RT / SK_THREADS;
                const int r = tid % RT, c0 = (tid / RT) * NC3;
                float dzr[AMAX];
#pragma unroll
                for (int j = 0; j < AMAX; ++j) dzr[j] = sDz[j * RS + r];
                const float dv = sDz[AMAX * RS + r];
#pragma unroll 4
                for (int q = 0; q < NC3; ++q) {
                    const int n = c0 + q;
                    const float h = sH3[n * RS + r];
                    float g;
                    if (n < HN) {
                        g = 0.f;
#pragma unroll
                        for (int j = 0; j < AMAX; ++j) if (j < A) g = fmaf(dzr[j], sP[L.wa + j * HN + n], g);
                    } else {
                        g = dv * sP[L.wc + (n - HN)];
                    }
                    sD3[n * RS + r] = g * (1.0f - h * h);
                }
            }
            // head weight gradients (from dz / dv and h3): each output owned by one thread
            for (int o = tid; o < (A + 1) * HN; o += SK_THREADS) {
                const int j = o / HN, k = o - j * HN;
                if (j < A) sG[L.wa + j * HN + k] += dotR<RT>(sDz + j * RS, sH3 + k * RS);
                else sG[L.wc + k] += dotR<RT>(sDz + AMAX * RS, sH3 + (HN + k) * RS);
            }
            if (tid < A) {
                sG[L.ba + tid] += sumR<RT>(sDz + tid * RS);
                if (CONT) sG[L.log_std + tid] += sumR<RT>(sDls + tid * RS);
            } else if (tid == AMAX) {
                sG[L.bc] += sumR<RT>(sDz + AMAX * RS);
            }
            __syncthreads();
            SK_MARK(7);
            // ---- backward (ppo.py:283) ----
            if (tid < 2 * HN) sG[L.b3 + tid] += sumR<RT>(sD3 + tid * RS);
            dgrad_layer<2 * HN, RT>(sD3, sP + L.w3, sH2, sD2);
            wgrad_layer<2 * HN, RT>(sD3, sH2, sG + L.w3);
            __syncthreads();
            SK_MARK(8);
            if (tid < HN) sG[L.b2 + tid] += sumR<RT>(sD2 + tid * RS);
            dgrad_layer<HN, RT>(sD2, sP + L.w2, sH1, sD1);
            wgrad_layer<HN, RT>(sD2, sH1, sG + L.w2);
            __syncthreads();
            SK_MARK(9);
            if (tid < HN) sG[L.b1 + tid] += sumR<RT>(sD1 + tid * RS);
            for (int o = tid; o < HN * D; o += SK_THREADS) {
                const int n = o / D, k = o - n * D;
                sG[L.w1 + o] += dotR<RT>(sD1 + n * RS, sX + k * RS);
            }
            __syncthreads();
        }

        SK_MARK(10);
        // next step's per-row scalars -> shared memory (their loads were issued at the top of this step)
        if (tid < my_rows && step + 1 < a.steps) {
            s_src[tid] = n_src;
            s_oldlp[tid] = n_oldlp;
            s_adv[tid] = n_src >= 0 ? (n_adv - adv_mean) / adv_denom : 0.f;
            s_ret[tid] = n_ret;
            if (!CONT) s_acti[tid] = n_acti;
            else {
#pragma unroll
                for (int j = 0; j < AMAX; ++j) if (j < A) s_actf[tid * AMAX + j] = n_actf[j];
            }
            src_next = src_next2;
        }
        // ---- loss sums of this CTA ----
        {
            const float p = warp_sum(l_pol), v = warp_sum(l_val), e = warp_sum(l_ent);
            if (lane == 0) { s_loss[0][warp] = p; s_loss[1][warp] = v; s_loss[2][warp] = e; }
            __syncthreads();
            if (tid < 3) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < SK_WARPS; ++w) s += s_loss[tid][w];
                sRed[tid] = s;
            }
        }
        cluster.sync();                                                    // (1) all eight partial gradients are complete
        SK_MARK(11);

        // ---- reduce-scatter over distributed shared memory: my slice of the summed gradient + its sum of squares ----
        double sq = 0.0;
        {
            const float* rg[SK_CL];
#pragma unroll
            for (int q = 0; q < SK_CL; ++q) rg[q] = cluster.map_shared_rank(sG, q);
            for (int i = s0 + 4 * tid; i < s1; i += 4 * SK_THREADS) {       // slices are multiples of 4 floats
                float4 v[SK_CL];
#pragma unroll
                for (int q = 0; q < SK_CL; ++q) v[q] = ld4(rg[q] + i);
                float4 g = v[0];
#pragma unroll
                for (int q = 1; q < SK_CL; ++q) { g.x += v[q].x; g.y += v[q].y; g.z += v[q].z; g.w += v[q].w; }    // fixed rank order
                *reinterpret_cast<float4*>(sGs + (i - s0)) = g;
                sq += (double)g.x * g.x + (double)g.y * g.y + (double)g.z * g.z + (double)g.w * g.w;
            }
        }
        sq = warp_sum_d(sq);
        if (lane == 0) s_wsum[warp] = sq;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SK_WARPS; ++w) s += s_wsum[w];
            *reinterpret_cast<double*>(sRed + 4) = s;
        }
        SK_MARK(12);
        cluster.sync();                                                    // (2) every slice's partial norm is published; partials consumed
        SK_MARK(13);
        for (int i = 4 * tid; i < total; i += 4 * SK_THREADS)              // ready for the next step (peers finished reading it)
            *reinterpret_cast<float4*>(sG + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0) {
            double s = 0.0;
            float lp = 0.f, lv = 0.f, le = 0.f;
#pragma unroll
            for (int q = 0; q < SK_CL; ++q) {
                const float* rr = cluster.map_shared_rank(sRed, q);
                s += *reinterpret_cast<const double*>(rr + 4);
                lp += rr[0]; lv += rr[1]; le += rr[2];
            }
            const float tot = (float)sqrt(s);
            float coef = a.max_norm / (tot + 1e-6f);                       // clip_grad.py:165-182
            s_coef = coef > 1.0f ? 1.0f : coef;
            if (rank == 0) {
                const float pol = lp * a.inv_m, val = 0.5f * lv * a.inv_m, ent = le * a.inv_m;     // ppo.py:270-274
                float* out = a.losses + (int64_t)step * 4;
                out[0] = pol; out[1] = val; out[2] = ent; out[3] = pol + a.vw * val + -a.beta * ent;   // ppo.py:276-280
                if (a.grad_norm_out && step == a.steps - 1) *a.grad_norm_out = tot;
            }
        }
        __syncthreads();
        // ---- clip + Adam on my slice (adam.py:531-547), new parameters broadcast into all eight copies ----
        const float coef = s_coef;
        const float bc2_sqrt = a.step_consts[2 * step], neg_step = a.step_consts[2 * step + 1];
        const bool last = step == a.steps - 1;
        {
            float* rp[SK_CL];
#pragma unroll
            for (int q = 0; q < SK_CL; ++q) rp[q] = cluster.map_shared_rank(sP, q);
            for (int i = s0 + 4 * tid; i < s1; i += 4 * SK_THREADS) {
                const float4 g4 = ld4(sGs + (i - s0)), m4 = ld4(sMs + (i - s0)), v4 = ld4(sVs + (i - s0)), p4 = ld4(sP + i);
                const float gs[4] = {g4.x, g4.y, g4.z, g4.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w},
                            ps[4] = {p4.x, p4.y, p4.z, p4.w};
                float go[4], mo[4], vo[4], po[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float gi = __fmul_rn(gs[c], coef);
                    float mi = ms[c], vi = vs[c];
                    mi = fmaf(a.w1, gi - mi, mi);
                    vi = __fadd_rn(__fmul_rn(vi, a.beta2), __fmul_rn(__fmul_rn(a.w2, gi), gi));
                    const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), a.eps);
                    po[c] = __fadd_rn(ps[c], __fmul_rn(neg_step, __fdiv_rn(mi, denom)));
                    go[c] = gi; mo[c] = mi; vo[c] = vi;
                }
                const float4 pn = make_float4(po[0], po[1], po[2], po[3]);
                *reinterpret_cast<float4*>(sMs + (i - s0)) = make_float4(mo[0], mo[1], mo[2], mo[3]);
                *reinterpret_cast<float4*>(sVs + (i - s0)) = make_float4(vo[0], vo[1], vo[2], vo[3]);
#pragma unroll
                for (int q = 0; q < SK_CL; ++q) *reinterpret_cast<float4*>(rp[q] + i) = pn;
                if (last) {
                    *reinterpret_cast<float4*>(a.P + i) = pn;
                    *reinterpret_cast<float4*>(a.M + i) = make_float4(mo[0], mo[1], mo[2], mo[3]);
                    *reinterpret_cast<float4*>(a.V + i) = make_float4(vo[0], vo[1], vo[2], vo[3]);
                    *reinterpret_cast<float4*>(a.G + i) = make_float4(go[0], go[1], go[2], go[3]);
                }
            }
        }
        SK_MARK(14);
        cluster.sync();                                                    // (3) all parameter copies updated
    }
}

size_t small_smem_bytes(const dppo_mlp_layout& L, int D, int RT, int CL)
{
    const int total = (int)L.total, RS = RT + 4;
    const int SL = ((total + CL - 1) / CL + 3) & ~3;
    const size_t fl = (size_t)2 * total + 3 * SL + (size_t)(D + 4 * HN + 2 * 2 * HN) * RS + RT * AMAX + RT + (size_t)(2 * AMAX + 1) * RS + 8;
    return fl * sizeof(float) + 64;
}

// shape of the launch for a minibatch of `rows`: 16-row tiles when a CTA of the 8-CTA cluster gets at most 16 rows, else 32-row
// tiles; 16 CTAs (non-portable cluster size, allowed on B200) once the 8-CTA split would give a CTA more than one 32-row tile
void small_plan(int64_t rows, int* RT, int* CL)
{
    *CL = rows > 8 * 32 ? 16 : 8;
    *RT = (rows + *CL - 1) / *CL <= 16 ? 16 : 32;
}

template <bool CONT, int RT, int CL>
int small_launch(dppo_ctx* ctx, const SmallArgs& a, cudaStream_t st)
{
    const size_t smem = small_smem_bytes(a.L, a.D, RT, CL);
    auto kern = small_update_kernel<CONT, RT, CL>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
        (CL > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess))
        DPPO_FAIL(ctx, "small_update: cudaFuncSetAttribute failed (%s)", cudaGetErrorString(cudaGetLastError()));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(CL); lc.blockDim = dim3(SK_THREADS); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    lc.attrs = attr; lc.numAttrs = 1;
    cudaLaunchKernelEx(&lc, kern, a);
    DPPO_CHECK_LAUNCH(ctx, "small_update_kernel");
    return 0;
}

}  // namespace

extern "C" int dppo_small_update_supported(const dppo_mlp_desc* d)
{
    if (!d) return 0;
    dppo_mlp_layout L;
    if (dppo_mlp_layout_compute(d, &L)) return 0;
    return d->hidden == HN && d->obs_dim >= 1 && d->obs_dim <= 64 && d->act_dim >= 1 && d->act_dim <= AMAX && L.total % 4 == 0 &&
           small_smem_bytes(L, d->obs_dim, 32, 8) <= 220 * 1024;
}

extern "C" int dppo_small_update(dppo_ctx* ctx, const dppo_mlp_desc* d, float* params, float* grads, float* exp_avg, float* exp_avg_sq,
                                 const float* obs, const void* actions, const float* old_log_probs, const float* adv, const float* returns,
                                 const double* adv_stats, const int32_t* idx, int64_t rows, int steps, const dppo_hyper* hy,
                                 const float* step_consts, float* losses, float* grad_norm_out, void* stream)
{
    if (!ctx) return 1;
    if (!d || !params || !grads || !exp_avg || !exp_avg_sq || !obs || !actions || !old_log_probs || !adv || !returns || !idx || !hy ||
        !step_consts || !losses)
        DPPO_FAIL(ctx, "small_update: null argument");
    if (!dppo_small_update_supported(d)) DPPO_FAIL(ctx, "small_update: unsupported network (hidden must be %d, obs_dim <= 64, actions <= %d)", HN, AMAX);
    int RT, CL;
    small_plan(rows, &RT, &CL);
    if (rows < 1 || rows > (int64_t)CL * ROWS_MAX || steps < 1) DPPO_FAIL(ctx, "small_update: bad shape rows=%lld steps=%d", (long long)rows, steps);
    if (hy->advantage_norm && (!adv_stats || hy->adv_count < 2)) DPPO_FAIL(ctx, "small_update: advantage_norm needs adv_stats and adv_count >= 2");
    SmallArgs a;
    dppo_mlp_layout_compute(d, &a.L);
    a.D = d->obs_dim; a.A = d->act_dim;
    a.P = params; a.G = grads; a.M = exp_avg; a.V = exp_avg_sq;
    a.obs = obs; a.actions = actions; a.old_logp = old_log_probs; a.adv = adv; a.ret = returns;
    a.adv_stats = adv_stats; a.adv_count = hy->adv_count; a.advantage_norm = hy->advantage_norm;
    a.idx = idx; a.rows = (int)rows; a.steps = steps; a.step_consts = step_consts;
    a.clip = hy->ppo_clip; a.vw = hy->value_loss_weight; a.beta = hy->entropy_beta;
    a.inv_m = 1.0f / (float)(hy->loss_denominator > 0 ? hy->loss_denominator : rows);
    a.max_norm = hy->grad_norm_clip; a.w1 = (float)(1.0 - hy->beta1); a.beta2 = (float)hy->beta2; a.w2 = (float)(1.0 - hy->beta2);
    a.eps = hy->adam_eps;
    a.losses = losses; a.grad_norm_out = grad_norm_out;
    a.prof = nullptr;
#ifdef DPPO_SMALL_PROFILE
    { static long long* p = nullptr; if (!p) { cudaMallocManaged(&p, 16 * sizeof(long long)); memset(p, 0, 128); } a.prof = p; g_small_prof = p; }
#endif
    cudaStream_t st = (cudaStream_t)stream;
#define SMALL_GO(CONT)                                                          \
    do {                                                                        \
        if (CL == 8 && RT == 16) return small_launch<CONT, 16, 8>(ctx, a, st);  \
        if (CL == 8) return small_launch<CONT, 32, 8>(ctx, a, st);              \
        return small_launch<CONT, 32, 16>(ctx, a, st);                          \
    } while (0)
    if (d->continuous) SMALL_GO(true);
    SMALL_GO(false);
#undef SMALL_GO
}
