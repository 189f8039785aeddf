// FP32 FMA-pipe GEMMs for the actor-critic MLP (diamond/ppo.py:91-96 forward, ppo.py:283 backward).
//
// fp32-exact arithmetic is required: single-pass TF32 misses the 1e-4 parameter bar by 70x
// (SURVEY.md §0.6), so these kernels run on the FFMA pipe with fp32 operands and accumulators.
// Three products cover every layer:
//   NT  C[M,N] = epi(A[M,K] * W[N,K]^T + b)            forward Linear (+tanh), optional row gather
//   NN  C[M,N] = (A[M,K] * W[K,N]) .* (1 - H^2)         dgrad through a Linear and the tanh before it
//   TN  P[s][N1,N2] = sum_{m in split s} A[m,N1] * B[m,N2]   split-K weight gradient (deterministic)
// Tiling: CTA tile BMxBN, k-step BK, each thread an (TM x TN) register tile split in two halves
// per dimension so that shared-memory fragment loads are 128-bit and broadcast within a warp;
// global->register prefetch of the next k-tile overlaps the FFMA block (2-stage pipeline).
#include "common.cuh"

namespace {

template <int BM, int BN, int BK>
struct Smem {
    float a[2][BK][BM + 4];
    float b[2][BK][BN + 4];
};

__device__ __forceinline__ float4 ld4_guard(const float* __restrict__ p, bool row_ok, int col, int ncols, bool vec_ok)
{
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!row_ok) return v;
    if (vec_ok && col + 3 < ncols) return __ldg(reinterpret_cast<const float4*>(p + col));
    if (col < ncols) v.x = __ldg(p + col);
    if (col + 1 < ncols) v.y = __ldg(p + col + 1);
    if (col + 2 < ncols) v.z = __ldg(p + col + 2);
    if (col + 3 < ncols) v.w = __ldg(p + col + 3);
    return v;
}

// 8x8 (or 4x4) outer-product update from one k-slice of the shared tiles.
template <int BM, int BN, int TM, int TN>
__device__ __forceinline__ void fma_block(const float (*__restrict__ as)[BM + 4], const float (*__restrict__ bs)[BN + 4],
                                          int BKc, int ty, int tx, float (&acc)[TM][TN])
{
#pragma unroll
    for (int k = 0; k < BKc; ++k) {
        float af[TM], bf[TN];
        if (TM == 8) {
            const float4 a0 = *reinterpret_cast<const float4*>(&as[k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&as[k][BM / 2 + ty * 4]);
            af[0] = a0.x; af[1] = a0.y; af[2] = a0.z; af[3] = a0.w;
            af[TM - 4] = a1.x; af[TM - 3] = a1.y; af[TM - 2] = a1.z; af[TM - 1] = a1.w;
        } else {
            const float4 a0 = *reinterpret_cast<const float4*>(&as[k][ty * 4]);
            af[0] = a0.x; af[1] = a0.y; af[2] = a0.z; af[3] = a0.w;
        }
        if (TN == 8) {
            const float4 b0 = *reinterpret_cast<const float4*>(&bs[k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&bs[k][BN / 2 + tx * 4]);
            bf[0] = b0.x; bf[1] = b0.y; bf[2] = b0.z; bf[3] = b0.w;
            bf[TN - 4] = b1.x; bf[TN - 3] = b1.y; bf[TN - 2] = b1.z; bf[TN - 1] = b1.w;
        } else {
            const float4 b0 = *reinterpret_cast<const float4*>(&bs[k][tx * 4]);
            bf[0] = b0.x; bf[1] = b0.y; bf[2] = b0.z; bf[3] = b0.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(af[i], bf[j], acc[i][j]);
    }
}

template <int BM, int TM>
__device__ __forceinline__ int tile_row(int ty, int i) { return TM == 8 ? (i < 4 ? ty * 4 + i : BM / 2 + ty * 4 + (i - 4)) : ty * 4 + i; }

// ---------------------------------------------------------------------------------------------
// NT / NN kernel.  MODE 0: B is [N,K] (k contiguous).  MODE 1: B is [K,N] (n contiguous).
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN, int MODE, int EPI>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), 2)
gemm_kernel(const float* __restrict__ A, int lda, const int32_t* __restrict__ a_rows, const float* __restrict__ B, int ldb,
            const float* __restrict__ bias, const float* __restrict__ Hact, int ldh, float* __restrict__ C, int ldc,
            float* __restrict__ colsum, int64_t M, int N, int K, int vecA, int vecB, int vecC, const int* __restrict__ m_dev)
{
    if (m_dev != nullptr) {                           // row count decided on the device: tiles past it exit at once
        const int64_t md = *m_dev;
        if (md < M) M = md;
        if ((int64_t)blockIdx.x * BM >= M) return;
    }
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;                      // threads along n
    __shared__ Smem<BM, BN, BK> sm;

    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // ---- global->register staging assignments ----
    // A tile [BM x BK], float4 along k: BM*BK/4 vectors over NT threads
    constexpr int A_VECS = BM * BK / 4, A_PER = (A_VECS + NT - 1) / NT;
    constexpr int AKQ = BK / 4;                       // float4 per tile row
    // B tile: MODE 0 [BN x BK] float4 along k; MODE 1 [BK x BN] float4 along n
    constexpr int B_VECS = BN * BK / 4, B_PER = (B_VECS + NT - 1) / NT;
    constexpr int BNQ = BN / 4;

    const float* a_ptr[A_PER];
    bool a_ok[A_PER];
#pragma unroll
    for (int v = 0; v < A_PER; ++v) {
        const int e = tid + v * NT;
        const int row = e / AKQ;
        const int64_t m = m0 + row;
        a_ok[v] = (e < A_VECS) && (m < M);
        int64_t src = m;
        if (a_ok[v] && a_rows) { src = a_rows[m]; if (src < 0) src = 0; }      // padding rows read row 0 (masked in the head kernel)
        a_ptr[v] = A + (a_ok[v] ? src : 0) * (int64_t)lda;
    }

    float4 ra[A_PER], rb[B_PER];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int v = 0; v < A_PER; ++v) {
            const int e = tid + v * NT;
            const int kq = e % AKQ;
            ra[v] = ld4_guard(a_ptr[v], a_ok[v], k0 + kq * 4, K, vecA);
        }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) {
            const int e = tid + v * NT;
            if (MODE == 0) {
                const int row = e / AKQ, kq = e % AKQ;
                const int n = n0 + row;
                rb[v] = ld4_guard(B + (int64_t)(n < N ? n : 0) * ldb, (e < B_VECS) && n < N, k0 + kq * 4, K, vecB);
            } else {
                const int krow = e / BNQ, nq = e % BNQ;
                const int k = k0 + krow;
                rb[v] = ld4_guard(B + (int64_t)(k < K ? k : 0) * ldb, (e < B_VECS) && k < K, n0 + nq * 4, N, vecB);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int v = 0; v < A_PER; ++v) {
            const int e = tid + v * NT;
            if (e < A_VECS) {
                const int row = e / AKQ, kq = e % AKQ;
                sm.a[buf][kq * 4 + 0][row] = ra[v].x;
                sm.a[buf][kq * 4 + 1][row] = ra[v].y;
                sm.a[buf][kq * 4 + 2][row] = ra[v].z;
                sm.a[buf][kq * 4 + 3][row] = ra[v].w;
            }
        }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) {
            const int e = tid + v * NT;
            if (e < B_VECS) {
                if (MODE == 0) {
                    const int row = e / AKQ, kq = e % AKQ;
                    sm.b[buf][kq * 4 + 0][row] = rb[v].x;
                    sm.b[buf][kq * 4 + 1][row] = rb[v].y;
                    sm.b[buf][kq * 4 + 2][row] = rb[v].z;
                    sm.b[buf][kq * 4 + 3][row] = rb[v].w;
                } else {
                    const int krow = e / BNQ, nq = e % BNQ;
                    *reinterpret_cast<float4*>(&sm.b[buf][krow][nq * 4]) = rb[v];
                }
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    for (int kt = 0; kt < nk; ++kt) {
        if (kt + 1 < nk) load_tiles((kt + 1) * BK);
        fma_block<BM, BN, TM, TN>(sm.a[buf], sm.b[buf], BK, ty, tx, acc);
        if (kt + 1 < nk) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---- epilogue ----
    float csum[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) csum[j] = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t m = m0 + tile_row<BM, TM>(ty, i);
        if (m >= M) continue;
#pragma unroll
        for (int jh = 0; jh < TN / 4; ++jh) {
            const int n = n0 + (TN == 8 ? (jh == 0 ? tx * 4 : BN / 2 + tx * 4) : tx * 4);
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x = acc[i][jh * 4 + j];
                const int nn = n + j;
                if (nn < N) {
                    if (EPI == DPPO_EPI_BIAS) x += __ldg(bias + nn);
                    if (EPI == DPPO_EPI_BIAS_TANH) x = tanhf(x + __ldg(bias + nn));
                    if (EPI == DPPO_EPI_TANH_BWD) {
                        const float h = __ldg(Hact + m * ldh + nn);
                        x = x * (1.0f - h * h);
                    }
                } else {
                    x = 0.f;
                }
                o[j] = x;
                csum[jh * 4 + j] += x;
            }
            float* dst = C + m * ldc + n;
            if (vecC && n + 3 < N) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < N) dst[j] = o[j];
            }
        }
    }
    if (EPI == DPPO_EPI_TANH_BWD && colsum != nullptr) {
        // column sums of this CTA's output tile -> bias-gradient partial (deterministic order)
        constexpr int TYN = BM / TM;
        float* red = &sm.a[0][0][0];                 // reuse: needs TYN*BN floats <= 2*BK*(BM+4)
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int nl = TN == 8 ? (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * 4 + j;
            red[ty * BN + nl] = csum[j];
        }
        __syncthreads();
        for (int nl = tid; nl < BN; nl += NT) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < TYN; ++r) s += red[r * BN + nl];
            if (n0 + nl < N) colsum[(int64_t)blockIdx.x * N + n0 + nl] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TN split-K weight gradient.  grid = (n2 tiles, n1 tiles, splits).
// ---------------------------------------------------------------------------------------------
template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN), 2)
wgrad_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb, const int32_t* __restrict__ b_rows,
             float* __restrict__ P, int64_t M, int N1, int N2, int64_t rows_per_split, int vecA, int vecB, int vecP)
{
    constexpr int NT = (BM / TM) * (BN / TN);
    constexpr int TXN = BN / TN;
    __shared__ Smem<BM, BN, BK> sm;
    const int tid = threadIdx.x;
    const int tx = tid % TXN, ty = tid / TXN;
    const int n2_0 = blockIdx.x * BN;
    const int n1_0 = blockIdx.y * BM;
    const int64_t r0 = (int64_t)blockIdx.z * rows_per_split;
    const int64_t r1 = (r0 + rows_per_split < M) ? r0 + rows_per_split : M;

    constexpr int AQ = BM / 4, BQ = BN / 4;
    constexpr int A_VECS = BK * AQ, A_PER = (A_VECS + NT - 1) / NT;
    constexpr int B_VECS = BK * BQ, B_PER = (B_VECS + NT - 1) / NT;
    float4 ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int64_t k0) {
#pragma unroll
        for (int v = 0; v < A_PER; ++v) {
            const int e = tid + v * NT;
            const int krow = e / AQ, q = e % AQ;
            const int64_t m = k0 + krow;
            const bool ok = (e < A_VECS) && m < r1;
            ra[v] = ld4_guard(A + (ok ? m : 0) * (int64_t)lda, ok, n1_0 + q * 4, N1, vecA);
        }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) {
            const int e = tid + v * NT;
            const int krow = e / BQ, q = e % BQ;
            const int64_t m = k0 + krow;
            const bool ok = (e < B_VECS) && m < r1;
            int64_t src = ok ? m : 0;
            if (ok && b_rows) { src = b_rows[m]; if (src < 0) src = 0; }
            rb[v] = ld4_guard(B + src * (int64_t)ldb, ok, n2_0 + q * 4, N2, vecB);
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int v = 0; v < A_PER; ++v) {
            const int e = tid + v * NT;
            if (e < A_VECS) *reinterpret_cast<float4*>(&sm.a[buf][e / AQ][(e % AQ) * 4]) = ra[v];
        }
#pragma unroll
        for (int v = 0; v < B_PER; ++v) {
            const int e = tid + v * NT;
            if (e < B_VECS) *reinterpret_cast<float4*>(&sm.b[buf][e / BQ][(e % BQ) * 4]) = rb[v];
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int nk = (int)((r1 - r0 + BK - 1) / BK);
    if (nk > 0) {
        load_tiles(r0);
        store_tiles(0);
        __syncthreads();
        int buf = 0;
        for (int kt = 0; kt < nk; ++kt) {
            if (kt + 1 < nk) load_tiles(r0 + (int64_t)(kt + 1) * BK);
            fma_block<BM, BN, TM, TN>(sm.a[buf], sm.b[buf], BK, ty, tx, acc);
            if (kt + 1 < nk) store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
    float* Ps = P + (int64_t)blockIdx.z * N1 * N2;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int n1 = n1_0 + tile_row<BM, TM>(ty, i);
        if (n1 >= N1) continue;
#pragma unroll
        for (int jh = 0; jh < TN / 4; ++jh) {
            const int n2 = n2_0 + (TN == 8 ? (jh == 0 ? tx * 4 : BN / 2 + tx * 4) : tx * 4);
            float* dst = Ps + (int64_t)n1 * N2 + n2;
            if (vecP && n2 + 3 < N2) {
                *reinterpret_cast<float4*>(dst) = make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n2 + j < N2) dst[j] = acc[i][jh * 4 + j];
            }
        }
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Large tile for dense layers, small tile when the problem is tiny (H=64 nets, 128-row minibatches).
constexpr int LBM = 128, LBN = 128, LBK = 16, LT = 8;
constexpr int SBM = 64, SBN = 64, SBK = 16, ST = 4;

inline bool use_large(int64_t M, int N) { return M >= 1024 && N >= 128; }

template <int MODE, int EPI>
int launch_gemm(dppo_ctx* ctx, const float* A, int lda, const int32_t* a_rows, const float* B, int ldb, const float* bias,
                const float* Hact, int ldh, float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st)
{
    if (M <= 0 || N <= 0 || K <= 0) DPPO_FAIL(ctx, "gemm: empty shape M=%lld N=%d K=%d", (long long)M, N, K);
    const int vecA = aligned16(A) && lda % 4 == 0;
    const int vecB = aligned16(B) && ldb % 4 == 0;
    const int vecC = aligned16(C) && ldc % 4 == 0;
    if (use_large(M, N)) {
        dim3 grid((unsigned)((M + LBM - 1) / LBM), (unsigned)((N + LBN - 1) / LBN));
        gemm_kernel<LBM, LBN, LBK, LT, LT, MODE, EPI><<<grid, 256, 0, st>>>(A, lda, a_rows, B, ldb, bias, Hact, ldh, C, ldc,
                                                                           colsum, M, N, K, vecA, vecB, vecC, ctx->rows_dev);
    } else {
        dim3 grid((unsigned)((M + SBM - 1) / SBM), (unsigned)((N + SBN - 1) / SBN));
        gemm_kernel<SBM, SBN, SBK, ST, ST, MODE, EPI><<<grid, 256, 0, st>>>(A, lda, a_rows, B, ldb, bias, Hact, ldh, C, ldc,
                                                                           colsum, M, N, K, vecA, vecB, vecC, ctx->rows_dev);
    }
    DPPO_CHECK_LAUNCH(ctx, "gemm_kernel");
    return 0;
}

}  // namespace

int dppo_gemm_nt(dppo_ctx* ctx, int epi, const float* A, int lda, const int32_t* a_rows, const float* B, int ldb,
                 const float* bias, float* C, int ldc, int64_t M, int N, int K, cudaStream_t st)
{
    if (epi == DPPO_EPI_BIAS_TANH)
        return launch_gemm<0, DPPO_EPI_BIAS_TANH>(ctx, A, lda, a_rows, B, ldb, bias, nullptr, 0, C, ldc, nullptr, M, N, K, st);
    return launch_gemm<0, DPPO_EPI_BIAS>(ctx, A, lda, a_rows, B, ldb, bias, nullptr, 0, C, ldc, nullptr, M, N, K, st);
}

int dppo_gemm_row_tiles(int64_t M, int N) { return (int)((M + (use_large(M, N) ? LBM : SBM) - 1) / (use_large(M, N) ? LBM : SBM)); }

int dppo_gemm_nn_tanh_bwd(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, const float* Hact, int ldh,
                          float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st)
{
    return launch_gemm<1, DPPO_EPI_TANH_BWD>(ctx, A, lda, nullptr, B, ldb, nullptr, Hact, ldh, C, ldc, colsum, M, N, K, st);
}

int dppo_gemm_nn(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int64_t M, int N, int K,
                 cudaStream_t st)
{
    return launch_gemm<1, DPPO_EPI_NONE>(ctx, A, lda, nullptr, B, ldb, nullptr, nullptr, 0, C, ldc, nullptr, M, N, K, st);
}

static inline bool wgrad_large(int N1, int N2) { return N1 >= 128 && N2 >= 128; }

int dppo_wgrad_splits(dppo_ctx* ctx, int64_t M, int N1, int N2)
{
    const int bm = wgrad_large(N1, N2) ? LBM : SBM, bn = wgrad_large(N1, N2) ? LBN : SBN;
    const int tiles = ((N1 + bm - 1) / bm) * ((N2 + bn - 1) / bn);
    int target = 2 * ctx->sm_count;                       // two resident CTAs per SM
    int splits = (target + tiles - 1) / tiles;
    const int64_t max_splits = (M + 63) / 64;             // at least 64 rows per split
    if (splits > max_splits) splits = (int)max_splits;
    if (splits < 1) splits = 1;
    return splits;
}

int dppo_wgrad(dppo_ctx* ctx, const float* A, int lda, const float* B, int ldb, const int32_t* b_rows, float* partials,
               int splits, int64_t M, int N1, int N2, cudaStream_t st)
{
    if (M <= 0 || N1 <= 0 || N2 <= 0 || splits <= 0) DPPO_FAIL(ctx, "wgrad: empty shape");
    const int vecA = aligned16(A) && lda % 4 == 0;
    const int vecB = aligned16(B) && ldb % 4 == 0;
    const int vecP = aligned16(partials) && N2 % 4 == 0 && ((int64_t)N1 * N2) % 4 == 0;
    int64_t rps = (M + splits - 1) / splits;
    rps = (rps + 15) / 16 * 16;
    if (wgrad_large(N1, N2)) {
        dim3 grid((N2 + LBN - 1) / LBN, (N1 + LBM - 1) / LBM, splits);
        wgrad_kernel<LBM, LBN, LBK, LT, LT><<<grid, 256, 0, st>>>(A, lda, B, ldb, b_rows, partials, M, N1, N2, rps, vecA, vecB, vecP);
    } else {
        dim3 grid((N2 + SBN - 1) / SBN, (N1 + SBM - 1) / SBM, splits);
        wgrad_kernel<SBM, SBN, SBK, ST, ST><<<grid, 256, 0, st>>>(A, lda, B, ldb, b_rows, partials, M, N1, N2, rps, vecA, vecB, vecP);
    }
    DPPO_CHECK_LAUNCH(ctx, "wgrad_kernel");
    return 0;
}
