// Warp-specialised, persistent 3xTF32 tcgen05 kernels of the wide (H >= 128) actor-critic MLP update
// (diamond/ppo.py:261 forward, :283 backward).  fp32 parity needs error-compensated TF32:
// every fp32 operand x is split into hi = rn_tf32(x) and lo = x - hi and each k-step issues
//   D += A_hi*B_lo ; D += A_lo*B_hi ; D += A_hi*B_hi          (fp32 accumulators in TMEM).
//
// Roles inside one 448-thread CTA (one CTA per SM):
//   warp 0      producer : one lane streams operand chunks into a shared-memory ring with TMA
//                          (cp.async.bulk.tensor for fp32 activations, cp.async.bulk for the pre-split weight images)
//   warp 1      MMA      : one lane issues tcgen05.mma.kind::tf32 and commits to mbarriers
//   warps 2-5   split    : turn the raw fp32 activation chunk that TMA delivered into its hi (in place) and lo images
//   warps 6-13  epilogue : tcgen05.ld the finished accumulator, apply the layer epilogue, store to HBM
// so loads, the hi/lo split, the MMAs and the epilogue of consecutive tiles all overlap.
//
// tc2_gemm_kernel  : C[M,N] = epi(A[M,K] * B^T), A fp32 row-major (K-major operand, SWIZZLE_64B, 16-wide k chunks),
//                    B = weight images from prep_weights_kernel (gemm_tc.cu).  Two 128x256 accumulators in TMEM.
// tc2_wgrad_kernel : dW[N1,N2] = sum_m D[m,N1] * H[m,N2] (split over row ranges, deterministic partials).  Both
//                    operands are read as they lie in HBM (row-major, so M/N-major for the MMA, SWIZZLE_128B_BASE32B).
#include <cuda.h>

#include "tc_common.cuh"
#include "gemm_tc.cuh"

using namespace tc;

namespace {

constexpr int THREADS = 448;
constexpr int W_PROD = 0, W_MMA = 1, W_SPLIT0 = 2, N_SPLIT = 4, W_EPI0 = 6, N_EPI = 8;
constexpr int SPLIT_THREADS = N_SPLIT * 32;
constexpr int MAX_STAGES = 6;

// ---- forward / dgrad GEMM -------------------------------------------------------------------------
constexpr int KC = 16;                     // k per chunk (one SWIZZLE_64B atom width)
constexpr int BM = 128;                    // rows per tile (UMMA M)
constexpr int A_IMG = BM * KC * 4;         // 8 KB: one A image (hi or lo) of a chunk
constexpr int G_STAGES = 4;
constexpr int MAXN = 512;

template <int EPI>
__global__ void __launch_bounds__(THREADS, 1)
tc2_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const unsigned char* __restrict__ Wimg, const float* __restrict__ bias,
                const float* __restrict__ Hact, int ldh, float* __restrict__ C, int ldc, float* __restrict__ colsum, int64_t M,
                int N, int K, int n_tile, int m_tiles)
{
    extern __shared__ unsigned char dyn_raw[];
    __shared__ __align__(8) uint64_t full[G_STAGES], ready[G_STAGES], empty[G_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(16) float s_bias[MAXN];

    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~(uintptr_t)1023);
    const int b_img = n_tile * KC * 4;
    const int stage_bytes = 2 * A_IMG + 2 * b_img;           // [A_hi | A_lo | B_hi | B_lo]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_tiles = N / n_tile;
    const int total = m_tiles * n_tiles;
    const int chunks = K / KC;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
#pragma unroll
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], N_EPI); }
        fence_mbar_init();
    }
    if (EPI == DPPO_EPI_BIAS_TANH)
        for (int i = tid; i < N; i += THREADS) s_bias[i] = bias[i];
    if (warp == W_MMA) tmem_alloc(&s_tmem, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == W_PROD) {
        if (lane == 0) {
            tma_prefetch_desc(&tmA);
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
                const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
                const unsigned char* wsrc = Wimg + (int64_t)n_blk * chunks * 2 * b_img;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_expect_tx(&full[s], (uint32_t)(A_IMG + 2 * b_img));
                    unsigned char* st = dyn + s * stage_bytes;
                    tma_load_2d(st, &tmA, c * KC, m_blk * BM, &full[s]);
                    bulk_copy_g2s(st + 2 * A_IMG, wsrc + (int64_t)c * 2 * b_img, 2u * (uint32_t)b_img, &full[s]);
                    if (++s == G_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(BM, n_tile, 0, 0);
            int s = 0;
            uint32_t ph = 0, it = 0;
            for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
                const uint32_t acc = it & 1;
                mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);          // epilogue has drained this accumulator
                fence_after();
                const uint32_t d = tmem + acc * 256;
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&full[s], ph);
                    mbar_wait(&ready[s], ph);
                    fence_after();
                    const uint32_t a_hi = smem_u32(dyn + s * stage_bytes), a_lo = a_hi + A_IMG;
                    const uint32_t b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + (uint32_t)b_img;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint32_t ko = ks * 32;                    // 8 tf32 = 32 bytes along K inside the swizzle atom
                        umma_tf32(d, desc_k_sw64(a_hi + ko), desc_k_sw64(b_lo + ko), idesc, (c | ks) != 0);
                        umma_tf32(d, desc_k_sw64(a_lo + ko), desc_k_sw64(b_hi + ko), idesc, 1u);
                        umma_tf32(d, desc_k_sw64(a_hi + ko), desc_k_sw64(b_hi + ko), idesc, 1u);
                    }
                    umma_commit(&empty[s]);                             // frees the stage once these MMAs have read it
                    if (++s == G_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
        __syncwarp();
    } else if (warp < W_EPI0) {
        // hi/lo split of the activation chunk TMA delivered (raw fp32 lands in the hi image)
        const int ct = tid - W_SPLIT0 * 32;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&full[s], ph);
                float4* hi = reinterpret_cast<float4*>(dyn + s * stage_bytes);
                float4* lo = hi + A_IMG / 16;
                float4 x[A_IMG / 16 / SPLIT_THREADS];
#pragma unroll
                for (int j = 0; j < A_IMG / 16 / SPLIT_THREADS; ++j) x[j] = hi[ct + j * SPLIT_THREADS];
#pragma unroll
                for (int j = 0; j < A_IMG / 16 / SPLIT_THREADS; ++j) {
                    const float4 l = split_tf32x4(x[j]);
                    hi[ct + j * SPLIT_THREADS] = x[j];
                    lo[ct + j * SPLIT_THREADS] = l;
                }
                fence_proxy_async();                                    // generic-proxy writes -> visible to the tensor core
                mbar_arrive(&ready[s]);
                if (++s == G_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // epilogue: warp -> TMEM lane quadrant (warp % 4) and column half
        const int q = warp & 3, half = (warp - W_EPI0) >> 2;
        const int cols_per_half = n_tile / 2;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const int n0 = n_blk * n_tile;
            const uint32_t acc = it & 1;
            mbar_wait(&tfull[acc], (it >> 1) & 1);
            fence_after();
            const int64_t m = (int64_t)m_blk * BM + q * 32 + lane;
            for (int cb = 0; cb < cols_per_half; cb += 32) {
                const int col = half * cols_per_half + cb;
                float v[32];
                tmem_ld32(tmem + acc * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
                const int n = n0 + col;
                if (EPI == DPPO_EPI_BIAS_TANH) {
                    if (m < M) {
                        float* dst = C + m * (int64_t)ldc + n;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = *reinterpret_cast<const float4*>(s_bias + n + j);
                            float4 o;
                            o.x = tanhf(v[j] + b.x); o.y = tanhf(v[j + 1] + b.y);
                            o.z = tanhf(v[j + 2] + b.z); o.w = tanhf(v[j + 3] + b.w);
                            *reinterpret_cast<float4*>(dst + j) = o;
                        }
                    }
                } else {
                    if (m < M) {
                        const float* hp = Hact + m * (int64_t)ldh + n;
                        float* dst = C + m * (int64_t)ldc + n;
                        float4 hh[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) hh[j] = __ldg(reinterpret_cast<const float4*>(hp + 4 * j));
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[4 * j] *= (1.0f - hh[j].x * hh[j].x); v[4 * j + 1] *= (1.0f - hh[j].y * hh[j].y);
                            v[4 * j + 2] *= (1.0f - hh[j].z * hh[j].z); v[4 * j + 3] *= (1.0f - hh[j].w * hh[j].w);
                            *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
                    if (colsum != nullptr) {
                        // warp transpose-reduce: afterwards lane l holds the sum over the warp's 32 rows of column l
#pragma unroll
                        for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
                            for (int i = 0; i < o; ++i) {
                                const bool up = lane & o;
                                const float send = up ? v[i] : v[i + o];
                                const float keep = up ? v[i + o] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                            }
                        }
                        colsum[((int64_t)m_blk * 4 + q) * N + n + lane] = v[0];
                    }
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ---- weight gradient -------------------------------------------------------------------------------
constexpr int WKC = 16;                    // rows (the contraction index) per chunk: two k8 MMA steps of two 4-row swizzle atoms
constexpr int GRP = WKC * 128;             // bytes of one 32-column group of a chunk

// grid = n1_blocks * splits.  CTA (n1_blk, split) accumulates rows [split*rows_per_split, +rows_per_split) of
//   partials[split][n1_blk*NACC*128 + i][j] = sum_m D[m, n1_0 + i] * H[m, j],   i < NACC*128, j < N2
// in NACC TMEM accumulators of N2 columns.  Stage layout: [D_hi | H_hi | D_lo | H_lo]; each operand is stored as
// 32-column groups of WKC rows x 128 B, exactly what a SWIZZLE_128B_ATOM_32B TMA box of 32 x WKC floats delivers.
template <int NACC>
__global__ void __launch_bounds__(THREADS, 1)
tc2_wgrad_kernel(const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmH, float* __restrict__ partials,
                 int64_t M, int N1, int N2, int n1_blocks, int rows_per_split, int stages)
{
    extern __shared__ unsigned char dyn_raw[];
    __shared__ __align__(8) uint64_t full[MAX_STAGES], ready[MAX_STAGES], empty[MAX_STAGES], tfull;
    __shared__ uint32_t s_tmem;

    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int gA = NACC * 4;
    const int gB = N2 / 32;
    constexpr int a_bytes = gA * GRP;
    const int b_bytes = gB * GRP;
    const int hi_bytes = a_bytes + b_bytes;
    const int stage_bytes = 2 * hi_bytes;
    const int n1_blk = blockIdx.x % n1_blocks, split = blockIdx.x / n1_blocks;
    const int n1_0 = n1_blk * NACC * 128;
    const int64_t r0 = (int64_t)split * rows_per_split;
    const int64_t r1 = r0 + rows_per_split < M ? r0 + rows_per_split : M;
    const int chunks = r1 > r0 ? (int)((r1 - r0 + WKC - 1) / WKC) : 0;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&full[s], 1); mbar_init(&ready[s], SPLIT_THREADS); mbar_init(&empty[s], 1); }
        mbar_init(&tfull, 1);
        fence_mbar_init();
    }
    if (warp == W_MMA) tmem_alloc(&s_tmem, 512);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = s_tmem;

    if (warp == W_PROD) {
        if (lane == 0) {
            tma_prefetch_desc(&tmD);
            tma_prefetch_desc(&tmH);
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)hi_bytes);
                unsigned char* st = dyn + s * stage_bytes;
                const int row = (int)(r0 + (int64_t)c * WKC);
#pragma unroll
                for (int g = 0; g < gA; ++g) tma_load_2d(st + g * GRP, &tmD, n1_0 + g * 32, row, &full[s]);
                for (int g = 0; g < gB; ++g) tma_load_2d(st + a_bytes + g * GRP, &tmH, g * 32, row, &full[s]);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == W_MMA) {
        if (lane == 0) {
            const uint32_t idesc = idesc_tf32(128, N2, 1, 1);
            int s = 0;
            uint32_t ph = 0;
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&full[s], ph);
                mbar_wait(&ready[s], ph);
                fence_after();
                const uint32_t base = smem_u32(dyn + s * stage_bytes);
#pragma unroll
                for (int j = 0; j < WKC / 8; ++j) {
                    const uint32_t b_hi = base + a_bytes + j * 1024, b_lo = b_hi + hi_bytes;
                    const uint64_t dbh = desc_mn_sw128_32b(b_hi, GRP), dbl = desc_mn_sw128_32b(b_lo, GRP);
#pragma unroll
                    for (int a = 0; a < NACC; ++a) {
                        const uint32_t a_hi = base + a * 4 * GRP + j * 1024, a_lo = a_hi + hi_bytes;
                        const uint64_t dah = desc_mn_sw128_32b(a_hi, GRP), dal = desc_mn_sw128_32b(a_lo, GRP);
                        const uint32_t d = tmem + (uint32_t)(a * N2);
                        umma_tf32(d, dah, dbl, idesc, (c | j) != 0);
                        umma_tf32(d, dal, dbh, idesc, 1u);
                        umma_tf32(d, dah, dbh, idesc, 1u);
                    }
                }
                umma_commit(&empty[s]);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
            umma_commit(&tfull);
        }
        __syncwarp();
    } else if (warp < W_EPI0) {
        const int ct = tid - W_SPLIT0 * 32;
        const int n4 = hi_bytes / 16;
        int s = 0;
        uint32_t ph = 0;
        for (int c = 0; c < chunks; ++c) {
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(dyn + s * stage_bytes);
            float4* lo = hi + n4;
            for (int i = ct; i < n4; i += 4 * SPLIT_THREADS) {
                float4 x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i + j * SPLIT_THREADS < n4) x[j] = hi[i + j * SPLIT_THREADS];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i + j * SPLIT_THREADS < n4) {
                        const float4 l = split_tf32x4(x[j]);
                        hi[i + j * SPLIT_THREADS] = x[j];
                        lo[i + j * SPLIT_THREADS] = l;
                    }
            }
            fence_proxy_async();
            mbar_arrive(&ready[s]);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else {
        const int q = warp & 3, half = (warp - W_EPI0) >> 2;
        const int cols_per_half = N2 / 2;
        if (chunks > 0) {
            mbar_wait(&tfull, 0);
            fence_after();
        }
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            float* dst = partials + ((int64_t)split * N1 + n1_0 + a * 128 + q * 32 + lane) * N2;
            for (int cb = 0; cb < cols_per_half; cb += 32) {
                const int col = half * cols_per_half + cb;
                float v[32];
                if (chunks > 0) {
                    tmem_ld32(tmem + (uint32_t)(a * N2) + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + col + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }

    fence_before();
    __syncthreads();
    if (warp == W_MMA) {
        fence_after();
        tmem_dealloc(tmem, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

bool dppo_make_tensor_map_2d(CUtensorMap* out, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
                             int swizzle)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = swizzle == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_128B : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- launchers -------------------------------------------------------------------------------------
bool dppo_tc2_gemm_supported(int64_t M, int N, int K)
{
    return M >= 1024 && M < (int64_t)1 << 31 && K % KC == 0 && (N % 256 == 0 || N == 128) && N <= MAXN;
}

int dppo_tc2_colsum_parts(int64_t M) { return (int)((M + BM - 1) / BM) * 4; }

int dppo_tc2_gemm(dppo_ctx* ctx, int epi, const float* A, int lda, const unsigned char* Wimg, const float* bias, const float* Hact,
                  int ldh, float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st)
{
    if (!dppo_tc2_gemm_supported(M, N, K)) DPPO_FAIL(ctx, "tc2_gemm: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    if (lda % 4 != 0 || ldc % 4 != 0 || !al16(A) || !al16(C) || !al16(Wimg) || (Hact && (!al16(Hact) || ldh % 4 != 0)))
        DPPO_FAIL(ctx, "tc2_gemm: operands must be 16-byte aligned with row pitches multiple of 4 floats");
    CUtensorMap tmA;
    if (!dppo_make_tensor_map_2d(&tmA, A, M, K, lda, KC, BM, 2)) DPPO_FAIL(ctx, "tc2_gemm: cuTensorMapEncodeTiled failed");
    const int n_tile = dppo_tc_n_tile(N);
    const int m_tiles = (int)((M + BM - 1) / BM);
    const int total = m_tiles * (N / n_tile);
    const size_t smem = (size_t)G_STAGES * (2 * A_IMG + 2 * n_tile * KC * 4) + 1024;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
    if (epi == DPPO_EPI_BIAS_TANH) {
        cudaFuncSetAttribute(tc2_gemm_kernel<DPPO_EPI_BIAS_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc2_gemm_kernel<DPPO_EPI_BIAS_TANH><<<grid, THREADS, smem, st>>>(tmA, Wimg, bias, nullptr, 0, C, ldc, nullptr, M, N, K, n_tile, m_tiles);
    } else if (epi == DPPO_EPI_TANH_BWD) {
        cudaFuncSetAttribute(tc2_gemm_kernel<DPPO_EPI_TANH_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc2_gemm_kernel<DPPO_EPI_TANH_BWD><<<grid, THREADS, smem, st>>>(tmA, Wimg, nullptr, Hact, ldh, C, ldc, colsum, M, N, K, n_tile, m_tiles);
    } else {
        DPPO_FAIL(ctx, "tc2_gemm: unknown epilogue %d", epi);
    }
    DPPO_CHECK_LAUNCH(ctx, "tc2_gemm_kernel");
    return 0;
}

bool dppo_tc2_wgrad_supported(int64_t M, int N1, int N2)
{
    return M >= 1024 && M < (int64_t)1 << 31 && N1 % 128 == 0 && N2 % 64 == 0 && N2 <= 256;
}

namespace {
struct WgradPlan { int nacc, n1_blocks, rows_per_split, splits, stages; size_t smem; };
WgradPlan wgrad_plan(int sm_count, int64_t M, int N1, int N2)
{
    WgradPlan p;
    p.nacc = (N1 % 256 == 0) ? 2 : 1;
    p.n1_blocks = N1 / (128 * p.nacc);
    int splits = sm_count / p.n1_blocks;
    if (splits < 1) splits = 1;
    int64_t rps = (M + splits - 1) / splits;
    rps = (rps + WKC - 1) / WKC * WKC;
    p.rows_per_split = (int)rps;
    p.splits = (int)((M + rps - 1) / rps);
    const int stage_bytes = 2 * (p.nacc * 4 + N2 / 32) * GRP;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > MAX_STAGES) p.stages = MAX_STAGES;
    p.smem = (size_t)p.stages * stage_bytes + 1024;
    return p;
}
}  // namespace

int dppo_tc2_wgrad_splits(dppo_ctx* ctx, int64_t M, int N1, int N2) { return wgrad_plan(ctx->sm_count, M, N1, N2).splits; }

int dppo_tc2_wgrad(dppo_ctx* ctx, const float* Dm, int ldd, const float* Hm, int ldh, float* partials, int splits, int64_t M, int N1,
                   int N2, cudaStream_t st)
{
    if (!dppo_tc2_wgrad_supported(M, N1, N2)) DPPO_FAIL(ctx, "tc2_wgrad: unsupported shape M=%lld N1=%d N2=%d", (long long)M, N1, N2);
    if (ldd % 4 != 0 || ldh % 4 != 0 || !al16(Dm) || !al16(Hm) || !al16(partials)) DPPO_FAIL(ctx, "tc2_wgrad: operands must be 16-byte aligned");
    const WgradPlan p = wgrad_plan(ctx->sm_count, M, N1, N2);
    if (p.splits != splits) DPPO_FAIL(ctx, "tc2_wgrad: caller sized the partials for %d splits, plan has %d", splits, p.splits);
    CUtensorMap tmD, tmH;
    if (!dppo_make_tensor_map_2d(&tmD, Dm, M, N1, ldd, 32, WKC, 4) || !dppo_make_tensor_map_2d(&tmH, Hm, M, N2, ldh, 32, WKC, 4))
        DPPO_FAIL(ctx, "tc2_wgrad: cuTensorMapEncodeTiled failed");
    const int grid = p.n1_blocks * p.splits;
    if (p.nacc == 2) {
        cudaFuncSetAttribute(tc2_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        tc2_wgrad_kernel<2><<<grid, THREADS, p.smem, st>>>(tmD, tmH, partials, M, N1, N2, p.n1_blocks, p.rows_per_split, p.stages);
    } else {
        cudaFuncSetAttribute(tc2_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
        tc2_wgrad_kernel<1><<<grid, THREADS, p.smem, st>>>(tmD, tmH, partials, M, N1, N2, p.n1_blocks, p.rows_per_split, p.stages);
    }
    DPPO_CHECK_LAUNCH(ctx, "tc2_wgrad_kernel");
    return 0;
}
