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

namespace {

constexpr int SK_CL = 8;            // CTAs per cluster
constexpr int SK_THREADS = 256;
constexpr int RT = 16;              // rows per tile
constexpr int HN = 64;              // hidden width this kernel is specialised for
constexpr int AMAX = 8;             // actions / action dims

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
};

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// out[n][r] = tanh(b[n] + sum_k in[k][r] * W[n][k]) for the 16 rows of a tile; thread = (row r, column group cg)
template <int N>
__device__ __forceinline__ void fwd_layer(const float* __restrict__ in, int K, const float* __restrict__ W, const float* __restrict__ b,
                                          float* __restrict__ out)
{
    constexpr int NC = N / 16;
    const int r = threadIdx.x & 15, c0 = (threadIdx.x >> 4) * NC;
    float acc[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) acc[j] = b[c0 + j];
    if ((K & 3) == 0) {
        for (int k = 0; k < K; k += 4) {
            const float a0 = in[k * RT + r], a1 = in[(k + 1) * RT + r], a2 = in[(k + 2) * RT + r], a3 = in[(k + 3) * RT + r];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 w = ld4(W + (c0 + j) * K + k);
                acc[j] = fmaf(a0, w.x, acc[j]); acc[j] = fmaf(a1, w.y, acc[j]); acc[j] = fmaf(a2, w.z, acc[j]); acc[j] = fmaf(a3, w.w, acc[j]);
            }
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float a = in[k * RT + r];
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[j] = fmaf(a, W[(c0 + j) * K + k], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) out[(c0 + j) * RT + r] = tanhf(acc[j]);
}

// din[k][r] = (sum_n dout[n][r] * W[n][k]) * (1 - hin[k][r]^2), K = 64 outputs (4 per thread), W row-major [N][64]
template <int N>
__device__ __forceinline__ void dgrad_layer(const float* __restrict__ dout, const float* __restrict__ W, const float* __restrict__ hin,
                                            float* __restrict__ din)
{
    const int r = threadIdx.x & 15, k0 = (threadIdx.x >> 4) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
        const float d = dout[n * RT + r];
        const float4 w = ld4(W + n * HN + k0);
        acc.x = fmaf(d, w.x, acc.x); acc.y = fmaf(d, w.y, acc.y); acc.z = fmaf(d, w.z, acc.z); acc.w = fmaf(d, w.w, acc.w);
    }
    const float h0 = hin[k0 * RT + r], h1 = hin[(k0 + 1) * RT + r], h2 = hin[(k0 + 2) * RT + r], h3 = hin[(k0 + 3) * RT + r];
    din[k0 * RT + r] = acc.x * (1.0f - h0 * h0);
    din[(k0 + 1) * RT + r] = acc.y * (1.0f - h1 * h1);
    din[(k0 + 2) * RT + r] = acc.z * (1.0f - h2 * h2);
    din[(k0 + 3) * RT + r] = acc.w * (1.0f - h3 * h3);
}

__device__ __forceinline__ float dot16(const float* __restrict__ a, const float* __restrict__ b)
{
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RT; q += 4) {
        const float4 x = ld4(a + q), y = ld4(b + q);
        s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s); s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
    }
    return s;
}
__device__ __forceinline__ float sum16(const float* __restrict__ a)
{
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < RT; q += 4) { const float4 x = ld4(a + q); s += (x.x + x.y) + (x.z + x.w); }
    return s;
}

// g[n][k] += sum_r dout[n][r] * hin[k][r] for a [N][64] weight: thread owns N*64/256 consecutive k of one n
template <int N>
__device__ __forceinline__ void wgrad_layer(const float* __restrict__ dout, const float* __restrict__ hin, float* __restrict__ g)
{
    constexpr int PER = N * HN / SK_THREADS;            // 32 (N = 128) or 16 (N = 64)
    constexpr int TPN = HN / PER;                       // threads per output row
    const int n = threadIdx.x / TPN, k0 = (threadIdx.x % TPN) * PER;
    const float* d = dout + n * RT;
    const float4 d0 = ld4(d), d1 = ld4(d + 4), d2 = ld4(d + 8), d3 = ld4(d + 12);
#pragma unroll 8
    for (int k = 0; k < PER; ++k) {
        const float* h = hin + (k0 + k) * RT;
        const float4 h0 = ld4(h), h1 = ld4(h + 4), h2 = ld4(h + 8), h3 = ld4(h + 12);
        float s = d0.x * h0.x;
        s = fmaf(d0.y, h0.y, s); s = fmaf(d0.z, h0.z, s); s = fmaf(d0.w, h0.w, s);
        s = fmaf(d1.x, h1.x, s); s = fmaf(d1.y, h1.y, s); s = fmaf(d1.z, h1.z, s); s = fmaf(d1.w, h1.w, s);
        s = fmaf(d2.x, h2.x, s); s = fmaf(d2.y, h2.y, s); s = fmaf(d2.z, h2.z, s); s = fmaf(d2.w, h2.w, s);
        s = fmaf(d3.x, h3.x, s); s = fmaf(d3.y, h3.y, s); s = fmaf(d3.z, h3.z, s); s = fmaf(d3.w, h3.w, s);
        g[n * HN + k0 + k] += s;
    }
}

template <bool CONT>
__global__ void __cluster_dims__(SK_CL, 1, 1) __launch_bounds__(SK_THREADS, 1)
small_update_kernel(const SmallArgs a)
{
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int tid = threadIdx.x;
    const dppo_mlp_layout& L = a.L;
    const int total = (int)L.total, D = a.D, A = a.A;
    const int SL = ((total + SK_CL - 1) / SK_CL + 3) & ~3;              // slice of the flat buffers owned by one CTA
    const int s0 = rank * SL, s1 = min(total, s0 + SL);

    float* sP = smem;                          // [total] parameters (identical in all CTAs)
    float* sG = sP + total;                    // [total] this CTA's gradient partial
    float* sGs = sG + total;                   // [SL] reduced gradient slice
    float* sMs = sGs + SL;                     // [SL] exp_avg slice
    float* sVs = sMs + SL;                     // [SL] exp_avg_sq slice
    float* sX = sVs + SL;                      // [D][RT]
    float* sH1 = sX + ((D * RT + 3) & ~3);     // [64][RT]
    float* sH2 = sH1 + HN * RT;
    float* sH3 = sH2 + HN * RT;                // [128][RT]
    float* sD3 = sH3 + 2 * HN * RT;
    float* sD2 = sD3 + 2 * HN * RT;
    float* sD1 = sD2 + HN * RT;
    float* sZ = sD1 + HN * RT;                 // [RT][AMAX] head outputs, then d(loss)/dz
    float* sVal = sZ + RT * AMAX;              // [RT] values, then d(loss)/dv
    float* sDls = sVal + RT;                   // [RT][AMAX] per-row d(loss)/d(log_std) (continuous)
    float* sRed = sDls + RT * AMAX;            // [16] block reductions: [0..2] loss sums, [4..5] norm partial (double)
    __shared__ int s_src[RT];
    __shared__ float s_oldlp[RT], s_adv[RT], s_ret[RT], s_actf[RT * AMAX];
    __shared__ int s_acti[RT];
    __shared__ float s_coef;
    __shared__ double s_wsum[SK_THREADS / 32];

    for (int i = tid; i < total; i += SK_THREADS) sP[i] = a.P[i];
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
    cluster.sync();

    for (int step = 0; step < a.steps; ++step) {
        const int32_t* idx = a.idx + (int64_t)step * a.rows;
        for (int i = tid; i < total; i += SK_THREADS) sG[i] = 0.f;
        float l_pol = 0.f, l_val = 0.f, l_ent = 0.f;                      // thread tid < RT: sums over its rows
        __syncthreads();

        for (int t0 = row_lo; t0 < row_hi; t0 += RT) {
            // ---- gather (ppo.py:261-272) ----
            if (tid < RT) {
                const int m = t0 + tid;
                const int src = m < row_hi ? idx[m] : -1;
                s_src[tid] = src;
                const bool live = src >= 0;
                s_oldlp[tid] = live ? a.old_logp[src] : 0.f;
                s_adv[tid] = live ? (a.adv[src] - adv_mean) / adv_denom : 0.f;
                s_ret[tid] = live ? a.ret[src] : 0.f;
                if (!CONT) s_acti[tid] = live ? reinterpret_cast<const int32_t*>(a.actions)[src] : 0;
            }
            __syncthreads();
            for (int i = tid; i < D * RT; i += SK_THREADS) {
                const int r = i / D, k = i - r * D;                      // consecutive threads read consecutive floats of a row
                const int src = s_src[r];
                sX[k * RT + r] = src >= 0 ? a.obs[(int64_t)src * D + k] : 0.f;
            }
            if (CONT) {
                for (int i = tid; i < RT * A; i += SK_THREADS) {
                    const int r = i / A, j = i - r * A;
                    const int src = s_src[r];
                    s_actf[r * AMAX + j] = src >= 0 ? reinterpret_cast<const float*>(a.actions)[(int64_t)src * A + j] : 0.f;
                }
            }
            __syncthreads();
            // ---- forward (ppo.py:91-96) ----
            fwd_layer<HN>(sX, D, sP + L.w1, sP + L.b1, sH1);
            __syncthreads();
            fwd_layer<HN>(sH1, HN, sP + L.w2, sP + L.b2, sH2);
            __syncthreads();
            fwd_layer<2 * HN>(sH2, HN, sP + L.w3, sP + L.b3, sH3);
            __syncthreads();
            {   // output heads: thread = (row, output o): o < A actor, o == A critic
                const int r = tid & 15, o = tid >> 4;
                if (o <= A) {
                    const float* w = o < A ? sP + L.wa + o * HN : sP + L.wc;
                    const float* h = o < A ? sH3 : sH3 + HN * RT;
                    float s = o < A ? sP[L.ba + o] : sP[L.bc];
                    for (int k = 0; k < HN; k += 4) {
                        const float4 w4 = ld4(w + k);
                        s = fmaf(h[k * RT + r], w4.x, s); s = fmaf(h[(k + 1) * RT + r], w4.y, s);
                        s = fmaf(h[(k + 2) * RT + r], w4.z, s); s = fmaf(h[(k + 3) * RT + r], w4.w, s);
                    }
                    if (o < A) sZ[r * AMAX + o] = s; else sVal[r] = s;
                }
            }
            __syncthreads();
            // ---- distribution, loss terms, d(loss)/d(head outputs) (ppo.py:264-280) ----
            if (tid < RT) {
                const int r = tid;
                const bool live = s_src[r] >= 0;
                float z[AMAX], dz[AMAX];
#pragma unroll
                for (int j = 0; j < AMAX; ++j) { z[j] = j < A ? sZ[r * AMAX + j] : 0.f; dz[j] = 0.f; }
                float new_lp, entropy;
                float aux1[AMAX], aux2[AMAX];                             // discrete: p, lsm; gaussian: diff, var
                if (!CONT) {
                    float mx = -CUDART_INF_F;
#pragma unroll
                    for (int j = 0; j < AMAX; ++j) if (j < A) mx = fmaxf(mx, z[j]);
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < AMAX; ++j) if (j < A) s += expf(z[j] - mx);
                    const float lse = mx + logf(s);
                    entropy = 0.f; new_lp = 0.f;
                    const int act = s_acti[r];
#pragma unroll
                    for (int j = 0; j < AMAX; ++j) {
                        if (j < A) {
                            aux2[j] = z[j] - lse; aux1[j] = expf(aux2[j]);
                            entropy -= aux1[j] * aux2[j];
                            if (j == act) new_lp = aux2[j];
                        }
                    }
                } else {
                    new_lp = 0.f; entropy = 0.f;
#pragma unroll
                    for (int j = 0; j < AMAX; ++j) {
                        if (j < A) {
                            const float ls = sP[L.log_std + j];
                            const float sigma = expf(ls), log_scale = logf(sigma);
                            aux2[j] = sigma * sigma; aux1[j] = s_actf[r * AMAX + j] - z[j];
                            new_lp += -(aux1[j] * aux1[j]) / (2.0f * aux2[j]) - log_scale - 0.91893853320467274178f;
                            entropy += 0.5f + 0.91893853320467274178f + log_scale;
                        }
                    }
                }
                // clipped surrogate (ppo.py:266-270) and the gradient autograd assigns
                const float adv = s_adv[r];
                const float ratio = expf(new_lp - s_oldlp[r]);
                const float lo = 1.0f - a.clip, hi = 1.0f + a.clip;
                const float u1 = -adv * ratio, u2 = -adv * fminf(fmaxf(ratio, lo), hi);
                const bool in_range = ratio >= lo && ratio <= hi;
                const float wsel = in_range ? 1.0f : (u1 > u2 ? 1.0f : (u1 == u2 ? 0.5f : 0.0f));
                const float dlogp = (-adv * wsel * a.inv_m) * ratio;
                const float verr = sVal[r] - s_ret[r];
                float dv = 0.f;
                if (live) {
#pragma unroll
                    for (int j = 0; j < AMAX; ++j) {
                        if (j < A) {
                            if (!CONT) dz[j] = dlogp * ((j == s_acti[r] ? 1.0f : 0.0f) - aux1[j]) + (a.beta * a.inv_m) * aux1[j] * (aux2[j] + entropy);
                            else dz[j] = dlogp * aux1[j] / aux2[j];
                        }
                    }
                    dv = a.vw * verr * a.inv_m;
                    l_pol += fmaxf(u1, u2); l_val += verr * verr; l_ent += entropy;
                }
#pragma unroll
                for (int j = 0; j < AMAX; ++j) {
                    sZ[r * AMAX + j] = dz[j];
                    if (CONT) sDls[r * AMAX + j] = (live && j < A) ? dlogp * (aux1[j] * aux1[j] / aux2[j] - 1.0f) - a.beta * a.inv_m : 0.f;
                }
                sVal[r] = dv;
            }
            __syncthreads();
            // ---- backward into the first head layers: d3[n][r] ----
            {
                const int r = tid & 15, c0 = (tid >> 4) * 8;
                float dzr[AMAX];
#pragma unroll
                for (int j = 0; j < AMAX; ++j) dzr[j] = sZ[r * AMAX + j];
                const float dv = sVal[r];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int n = c0 + q;
                    const float h = sH3[n * RT + r];
                    float g;
                    if (n < HN) {
                        g = 0.f;
#pragma unroll
                        for (int j = 0; j < AMAX; ++j) if (j < A) g = fmaf(dzr[j], sP[L.wa + j * HN + n], g);
                    } else {
                        g = dv * sP[L.wc + (n - HN)];
                    }
                    sD3[n * RT + r] = g * (1.0f - h * h);
                }
            }
            // head weight gradients (from dz / dv and h3): each output owned by one thread
            for (int o = tid; o < (A + 1) * HN; o += SK_THREADS) {
                const int j = o / HN, k = o - j * HN;
                float s = 0.f;
                if (j < A) {
#pragma unroll
                    for (int r = 0; r < RT; ++r) s = fmaf(sZ[r * AMAX + j], sH3[k * RT + r], s);
                    sG[L.wa + j * HN + k] += s;
                } else {
                    s = dot16(sVal, sH3 + (HN + k) * RT);
                    sG[L.wc + k] += s;
                }
            }
            if (tid < A) {
                float s = 0.f, sl = 0.f;
#pragma unroll
                for (int r = 0; r < RT; ++r) { s += sZ[r * AMAX + tid]; if (CONT) sl += sDls[r * AMAX + tid]; }
                sG[L.ba + tid] += s;
                if (CONT) sG[L.log_std + tid] += sl;
            } else if (tid == AMAX) {
                sG[L.bc] += sum16(sVal);
            }
            __syncthreads();
            // ---- backward (ppo.py:283) ----
            if (tid < 2 * HN) sG[L.b3 + tid] += sum16(sD3 + tid * RT);
            dgrad_layer<2 * HN>(sD3, sP + L.w3, sH2, sD2);
            wgrad_layer<2 * HN>(sD3, sH2, sG + L.w3);
            __syncthreads();
            if (tid < HN) sG[L.b2 + tid] += sum16(sD2 + tid * RT);
            dgrad_layer<HN>(sD2, sP + L.w2, sH1, sD1);
            wgrad_layer<HN>(sD2, sH1, sG + L.w2);
            __syncthreads();
            if (tid < HN) sG[L.b1 + tid] += sum16(sD1 + tid * RT);
            for (int o = tid; o < HN * D; o += SK_THREADS) {
                const int n = o / D, k = o - n * D;
                sG[L.w1 + o] += dot16(sD1 + n * RT, sX + k * RT);
            }
            __syncthreads();
        }

        // ---- loss sums of this CTA ----
        if (tid < 32) {
            float p = tid < RT ? l_pol : 0.f, v = tid < RT ? l_val : 0.f, e = tid < RT ? l_ent : 0.f;
            p = warp_sum(p); v = warp_sum(v); e = warp_sum(e);
            if (tid == 0) { sRed[0] = p; sRed[1] = v; sRed[2] = e; }
        }
        cluster.sync();                                                    // (1) all eight partial gradients are complete

        // ---- reduce-scatter over distributed shared memory: my slice of the summed gradient + its sum of squares ----
        double sq = 0.0;
        for (int i = s0 + tid; i < s1; i += SK_THREADS) {
            float g = 0.f;
#pragma unroll
            for (int q = 0; q < SK_CL; ++q) g += cluster.map_shared_rank(sG, q)[i];          // fixed rank order
            sGs[i - s0] = g;
            sq += (double)g * (double)g;
        }
        sq = warp_sum_d(sq);
        if ((tid & 31) == 0) s_wsum[tid >> 5] = sq;
        __syncthreads();
        if (tid == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SK_THREADS / 32; ++w) s += s_wsum[w];
            *reinterpret_cast<double*>(sRed + 4) = s;
        }
        cluster.sync();                                                    // (2) every slice's partial norm is published
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
        for (int i = s0 + tid; i < s1; i += SK_THREADS) {
            const float gi = __fmul_rn(sGs[i - s0], coef);
            float mi = sMs[i - s0], vi = sVs[i - s0];
            mi = fmaf(a.w1, gi - mi, mi);
            vi = __fadd_rn(__fmul_rn(vi, a.beta2), __fmul_rn(__fmul_rn(a.w2, gi), gi));
            const float denom = __fadd_rn(__fdiv_rn(sqrtf(vi), bc2_sqrt), a.eps);
            const float pn = __fadd_rn(sP[i], __fmul_rn(neg_step, __fdiv_rn(mi, denom)));
            sMs[i - s0] = mi; sVs[i - s0] = vi;
#pragma unroll
            for (int q = 0; q < SK_CL; ++q) cluster.map_shared_rank(sP, q)[i] = pn;
            if (last) { a.P[i] = pn; a.M[i] = mi; a.V[i] = vi; a.G[i] = gi; }
        }
        cluster.sync();                                                    // (3) all parameter copies updated; partials may be zeroed
    }
}

size_t small_smem_bytes(const dppo_mlp_layout& L, int D)
{
    const int total = (int)L.total;
    const int SL = ((total + SK_CL - 1) / SK_CL + 3) & ~3;
    size_t fl = (size_t)2 * total + 3 * SL + ((D * RT + 3) & ~3) + (size_t)(HN * 3 + 2 * HN * 2 + HN) * RT + RT * AMAX + RT + RT * AMAX + 16;
    return fl * sizeof(float) + 64;
}

}  // namespace

extern "C" int dppo_small_update_supported(const dppo_mlp_desc* d)
{
    if (!d) return 0;
    dppo_mlp_layout L;
    if (dppo_mlp_layout_compute(d, &L)) return 0;
    return d->hidden == HN && d->obs_dim >= 1 && d->obs_dim <= 64 && d->act_dim >= 1 && d->act_dim <= AMAX && L.total % 4 == 0 &&
           small_smem_bytes(L, d->obs_dim) <= 220 * 1024;
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
    if (rows < 1 || rows > (1 << 24) || steps < 1) DPPO_FAIL(ctx, "small_update: bad shape rows=%lld steps=%d", (long long)rows, steps);
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
    const size_t smem = small_smem_bytes(a.L, a.D);
    cudaStream_t st = (cudaStream_t)stream;
    if (d->continuous) {
        cudaFuncSetAttribute(small_update_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        small_update_kernel<true><<<SK_CL, SK_THREADS, smem, st>>>(a);
    } else {
        cudaFuncSetAttribute(small_update_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        small_update_kernel<false><<<SK_CL, SK_THREADS, smem, st>>>(a);
    }
    DPPO_CHECK_LAUNCH(ctx, "small_update_kernel");
    return 0;
}
