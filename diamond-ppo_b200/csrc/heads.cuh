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
};

int head_partial_floats(int H, int A);
int head_train_blocks(dppo_ctx* ctx, int64_t M);
int launch_head_train_kernel(dppo_ctx* ctx, const HeadTrainArgs& a, int continuous, int blocks, cudaStream_t st);
int launch_head_eval(dppo_ctx* ctx, const float* ha, const float* hc, int ld, const float* wa, const float* ba, const float* wc,
                     const float* bc, float* head_out, float* values, int64_t rows, int H, int A, cudaStream_t st);
