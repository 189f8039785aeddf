#pragma once
#include "common.cuh"

constexpr int HEAD_WARPS = 8;

struct HeadTrainArgs {
    const float* h3;          // [M, 2H] tanh activations of actor_head.0 | critic_head.0
    float* d3;                // [M, 2H] gradient w.r.t. their pre-activations (output)
    const float *wa, *ba, *wc, *bc, *log_std;
    const int32_t* idx;       // minibatch row -> flat sample index (may be null)
    const int32_t* actions_i; // [B]     (discrete)
    const float* actions_f;   // [B, A]  (continuous)
    const float *old_logp, *adv, *ret;
    const double* adv_stats;
    int64_t adv_count;
    int advantage_norm;
    int64_t M;
    int H, A;
    float clip, vw, beta, inv_m;
    float* partials;          // [blocks, partial_stride]
    int partial_stride;
    int rev;                  // role-split kernel: rows are visited from the last to the first (L2 reuse, gemm_tc3.cu dppo_tc3_gemm)
    int h3_first;             // role-split kernel: h3 (never read again) is loaded with the L2 evict-first hint
    int keep_d3;              // role-split kernel: plain instead of streaming (evict-first) stores of d3, which the next launch reads
};

// Layout of one head-kernel partial: [dWa A*H | dba A | dWc H | dbc 1 | dlog_std A | db3 2H | losses 4], every block starting on a
// multiple of 4 floats so that the gradient reduction reads all of them as float4 (an unaligned db3 block took the scalar path
// and was the long pole of grad_reduce_kernel).
struct HeadOffsets { int dba, dwc, dbc, dls, b3, loss, total; };
__host__ __device__ inline HeadOffsets head_offsets(int H, int A)
{
    auto up4 = [](int x) { return (x + 3) / 4 * 4; };
    HeadOffsets o;
    o.dba = up4(A * H);
    o.dwc = up4(o.dba + A);
    o.dbc = up4(o.dwc + H);
    o.dls = up4(o.dbc + 1);
    o.b3 = up4(o.dls + A);
    o.loss = up4(o.b3 + 2 * H);
    o.total = o.loss + 4;
    return o;
}
int head_partial_floats(int H, int A);
int head_train_blocks(dppo_ctx* ctx, int64_t M, int H, int A);
int launch_head_train_kernel(dppo_ctx* ctx, const HeadTrainArgs& a, int continuous, int blocks, cudaStream_t st);
int launch_head_eval(dppo_ctx* ctx, const float* ha, const float* hc, int ld, const float* wa, const float* ba, const float* wc,
                     const float* bc, float* head_out, float* values, int64_t rows, int H, int A, int rev, cudaStream_t st);
