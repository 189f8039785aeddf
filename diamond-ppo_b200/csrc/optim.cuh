#pragma once
#include "common.cuh"

constexpr int DPPO_MAX_SEGS = 16;
struct GradSeg {
    int64_t dst;          // offset in the flat gradient
    int64_t count;        // floats
    const float* src;     // first partial
    int64_t stride;       // floats between partials
    int32_t nparts;
    int32_t pad;
};
struct GradSegTable {
    GradSeg seg[DPPO_MAX_SEGS];
    int32_t nseg;
};

// grads[0..total) = sum of partials per segment (zero outside any segment); also reduces the head
// kernel's loss partials into losses[0..3].
// sumsq_out (optional): grad_reduce_blocks(ctx, total) fp64 partial sums of squares of the assembled gradient
int grad_reduce_blocks(dppo_ctx* ctx, int64_t total);
int launch_grad_reduce(dppo_ctx* ctx, const GradSegTable& tab, float* grads, int64_t total, const float* loss_partials,
                       int loss_nparts, int64_t loss_stride, float vw, float beta, float inv_m, float* losses, double* sumsq_out,
                       cudaStream_t st);

// clip_grad_norm_ + Adam from `nparts` fp64 partial sums of squares of the (already summed) gradient
int launch_clip_adam(dppo_ctx* ctx, float* params, float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const double* partials,
                     int nparts, const dppo_hyper* h, float* grad_norm_out, cudaStream_t st);
