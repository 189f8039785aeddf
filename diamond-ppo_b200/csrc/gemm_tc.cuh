#pragma once
#include "common.cuh"

// 3xTF32 tcgen05 GEMMs (gemm_tc.cu)
bool dppo_tc_supported(int64_t M, int N, int K);
int dppo_tc_n_tile(int N);
int64_t dppo_tc_image_bytes(int N, int K);        // bytes of the hi/lo weight images of an [N, K] B operand
int dppo_tc_prep_weights(dppo_ctx* ctx, const float* W, int rows_w, int cols_w, int transpose, unsigned char* img, cudaStream_t st);
int dppo_tc_gemm(dppo_ctx* ctx, int epi, const float* A, int lda, const int32_t* a_rows, const unsigned char* Wimg, const float* bias,
                 const float* Hact, int ldh, float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st);
