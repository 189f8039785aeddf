// Warp-specialised 3xTF32 tcgen05 weight-gradient kernel of the wide (H >= 128) actor-critic MLP update (diamond/ppo.py:283)
// and the host-side TMA tensor-map helper shared by the tensor-core kernels.  fp32 parity needs error-compensated TF32
// (SURVEY.md 0.6): every fp32 operand x is split into hi (its top 19 bits) and lo = rn_tf32(x - hi) and each k-step issues
//   D += A_hi*B_lo ; D += A_lo*B_hi ; D += A_hi*B_hi          (fp32 accumulators in TMEM).
//
// Roles inside one 448-thread CTA (one CTA per SM):
//   warp 0      producer : one lane streams operand chunks into a shared-memory ring with TMA (cp.async.bulk.tensor)
//   warp 1      MMA      : one lane issues tcgen05.mma.kind::tf32 and commits to mbarriers
//   warps 2-5   split    : write the lo image of the raw fp32 chunk that TMA delivered (the raw chunk is the hi image)
//   warps 6-13  epilogue : tcgen05.ld the finished accumulator, staged TMA store of the partial
//
// tc2_wgrad_kernel : dW[N1,N2] = sum_m D[m,N1] * H[m,N2] (split over row ranges, deterministic partials).  Both
//                    operands are read as they lie in HBM (row-major, so M/N-major for the MMA, SWIZZLE_128B_BASE32B).
// (The forward / dgrad GEMMs live in gemm_tc3.cu.)

#include "tc_common.cuh"
#include "gemm_tc.cuh"

using namespace tc;

namespace {

constexpr int THREADS = 448;
constexpr int W_PROD = 0, W_MMA = 1, W_SPLIT0 = 2, N_SPLIT = 4, W_EPI0 = 6;
constexpr int SPLIT_THREADS = N_SPLIT * 32;
constexpr int MAX_STAGES = 6;


// ---- weight gradient -------------------------------------------------------------------------------
constexpr int WKC = 16;                    // rows (the contraction index) per chunk: two k8 MMA steps of two 4-row swizzle atoms
constexpr int GRP = WKC * 128;             // bytes of one 32-column group of a chunk

// grid = n1_blocks * splits.  CTA (n1_blk, split) accumulates rows [split*rows_per_split, +rows_per_split) of
//   partials[split][n1_blk*NACC*128 + i][j] = sum_m D[m, n1_0 + i] * H[m, j],   i < NACC*128, j < N2
// in NACC TMEM accumulators of N2 columns.  Stage layout: [D_hi | H_hi | D_lo | H_lo]; each operand is stored as
// 32-column groups of WKC rows x 128 B, exactly what a SWIZZLE_128B_ATOM_32B TMA box of 32 x WKC floats delivers.
// Up to three weight gradients share one launch (WgradJobs): CTA ranges [cta_begin, cta_begin + n1_blocks * splits) belong to
// job j, and the host plan (wgrad_multi_plan) sizes the row ranges so that every CTA of the launch carries the same tensor
// work -- one wave over all SMs instead of one wave per gradient, and ~2.6x fewer row-range partials to reduce.
struct WgradJob {
    CUtensorMap tmD, tmH, tmP;
    int N1, N2, n1_blocks, rows_per_split, cta_begin, cta_count;
};
struct WgradJobs {
    WgradJob j[3];
    int n;
    int smem_bytes;            // dynamic shared memory of the launch (each job derives its own stage count from it)
    int64_t M;
    int rn_hi;                 // A/B: store a round-to-nearest hi image instead of using the raw chunk as hi
    int prefetch;              // L2 prefetch of the operand chunks PF ahead (tc_prefetch bit 2)
    int in_first;              // operands are read with the L2 evict-first hint (row_sweep bit 4): every row is read once per n1 block
};

template <int NACC>
__global__ void __launch_bounds__(THREADS, 1)
tc2_wgrad_kernel(const __grid_constant__ WgradJobs jobs)
{
    extern __shared__ unsigned char dyn_raw[];
    __shared__ __align__(8) uint64_t full[MAX_STAGES], ready[MAX_STAGES], empty[MAX_STAGES], tfull;
    __shared__ uint32_t s_tmem;

    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int ji = 0;
    while (ji + 1 < jobs.n && (int)blockIdx.x >= jobs.j[ji].cta_begin + jobs.j[ji].cta_count) ++ji;
    const WgradJob& job = jobs.j[ji];
    const CUtensorMap& tmD = job.tmD;
    const CUtensorMap& tmH = job.tmH;
    const CUtensorMap& tmP = job.tmP;
    const int64_t M = jobs.M;
    const int N1 = job.N1, N2 = job.N2, n1_blocks = job.n1_blocks, rows_per_split = job.rows_per_split;
    const int local_cta = (int)blockIdx.x - job.cta_begin;
    constexpr int gA = NACC * 4;
    const int gB = N2 / 32;
    constexpr int a_bytes = gA * GRP;
    const int b_bytes = gB * GRP;
    const int hi_bytes = a_bytes + b_bytes;
    const int stage_bytes = 2 * hi_bytes;
    int stages = (jobs.smem_bytes - 1024) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const int n1_blk = local_cta % n1_blocks, split = local_cta / n1_blocks;
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
    DPPO_PDL_ENTER();                                    // set-up done; global memory only after the predecessor grid completed

    if (warp == W_PROD) {
        if (lane == 0) {
            tma_prefetch_desc(&tmD);
            tma_prefetch_desc(&tmH);
            int s = 0;
            uint32_t ph = 0;
            constexpr int PF = 8;                                       // chunks of L2 prefetch distance
            const bool inf = jobs.in_first != 0;
            const uint64_t pol = l2_policy_evict_first();
            const bool pf_on = jobs.prefetch != 0;
            for (int c = 0; pf_on && c < PF && c < chunks; ++c) {
                const int row = (int)(r0 + (int64_t)c * WKC);
                if (inf) {
                    for (int g = 0; g < gA; ++g) tma_prefetch_2d_hint(&tmD, n1_0 + g * 32, row, pol);
                    for (int g = 0; g < gB; ++g) tma_prefetch_2d_hint(&tmH, g * 32, row, pol);
                } else {
                    for (int g = 0; g < gA; ++g) tma_prefetch_2d(&tmD, n1_0 + g * 32, row);
                    for (int g = 0; g < gB; ++g) tma_prefetch_2d(&tmH, g * 32, row);
                }
            }
            for (int c = 0; c < chunks; ++c) {
                if (pf_on && c + PF < chunks) {
                    const int prow = (int)(r0 + (int64_t)(c + PF) * WKC);
                    if (inf) {
                        for (int g = 0; g < gA; ++g) tma_prefetch_2d_hint(&tmD, n1_0 + g * 32, prow, pol);
                        for (int g = 0; g < gB; ++g) tma_prefetch_2d_hint(&tmH, g * 32, prow, pol);
                    } else {
                        for (int g = 0; g < gA; ++g) tma_prefetch_2d(&tmD, n1_0 + g * 32, prow);
                        for (int g = 0; g < gB; ++g) tma_prefetch_2d(&tmH, g * 32, prow);
                    }
                }
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], (uint32_t)hi_bytes);
                unsigned char* st = dyn + s * stage_bytes;
                const int row = (int)(r0 + (int64_t)c * WKC);
                if (inf) {
#pragma unroll
                    for (int g = 0; g < gA; ++g) tma_load_2d_hint(st + g * GRP, &tmD, n1_0 + g * 32, row, &full[s], pol);
                    for (int g = 0; g < gB; ++g) tma_load_2d_hint(st + a_bytes + g * GRP, &tmH, g * 32, row, &full[s], pol);
                } else {
#pragma unroll
                    for (int g = 0; g < gA; ++g) tma_load_2d(st + g * GRP, &tmD, n1_0 + g * 32, row, &full[s]);
                    for (int g = 0; g < gB; ++g) tma_load_2d(st + a_bytes + g * GRP, &tmH, g * 32, row, &full[s]);
                }
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
            const uint32_t hi = smem_u32(dyn + s * stage_bytes);
            for (int i = ct; i < n4; i += 4 * SPLIT_THREADS) {
                float4 x[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (i + j * SPLIT_THREADS < n4) x[j] = lds128(hi + (i + j * SPLIT_THREADS) * 16);
                if (jobs.rn_hi) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i + j * SPLIT_THREADS < n4) {
                            const float4 l = split_tf32x4(x[j]);
                            sts128(hi + (i + j * SPLIT_THREADS) * 16, x[j]);
                            sts128(hi + hi_bytes + (i + j * SPLIT_THREADS) * 16, l);
                        }
                } else {                                                // the raw chunk is the hi image; only lo is written
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i + j * SPLIT_THREADS < n4) sts128(hi + hi_bytes + (i + j * SPLIT_THREADS) * 16, lo_of_trunc_x4(x[j]));
                }
            }
            fence_proxy_async();
            mbar_arrive(&ready[s]);
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else {
        // partials leave through double-buffered 32x32 staging blocks (the operand ring is free once the MMAs are done)
        // and TMA stores into the [splits * N1, N2] partial matrix
        const int ew = warp - W_EPI0;
        const int q = warp & 3, half = ew >> 2;
        const int cols_per_half = N2 / 2;
        if (chunks > 0) {
            mbar_wait(&tfull, 0);
            fence_after();
        }
        const uint32_t stg = smem_u32(dyn) + (uint32_t)ew * 2 * 4096;
        const uint32_t row_off = (uint32_t)lane * 128, sw = (uint32_t)(lane & 7);
        uint32_t blk = 0;
        if (lane == 0) tma_prefetch_desc(&tmP);
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            const int prow = split * N1 + n1_0 + a * 128 + q * 32;
            for (int cb = 0; cb < cols_per_half; cb += 32, ++blk) {
                const int col = half * cols_per_half + cb;
                float v[32];
                if (chunks > 0) {
                    tmem_ld32(tmem + (uint32_t)(a * N2) + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
                const uint32_t buf = stg + (blk & 1) * 4096;
                if (lane == 0) bulk_wait_read<1>();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts128(buf + row_off + ((((uint32_t)j) ^ sw) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmP, col, prow, buf);
                    bulk_commit();
                }
            }
        }
        if (lane == 0) bulk_wait<0>();
        __syncwarp();
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


bool dppo_tc2_wgrad_supported(int64_t M, int N1, int N2)
{
    return M >= 1024 && M < (int64_t)1 << 31 && N1 % 128 == 0 && N2 % 64 == 0 && N2 <= 256;
}

namespace {
struct WgradPlan { int nacc, n1_blocks, rows_per_split, splits; };

// measured cycles of one tcgen05.mma (M = 128, K = 8 tf32, shared-memory operands) as a function of N (dppo_tc_mma_probe)
double mma_clk(int n2) { return n2 >= 256 ? 171.0 : n2 >= 128 ? 107.0 + (n2 - 128) * (64.0 / 128.0) : 96.0 + (n2 - 64) * (11.0 / 64.0); }

int stage_bytes_of(int nacc, int N2) { return 2 * (nacc * 4 + N2 / 32) * GRP; }

// Row-range plan of `n` weight gradients sharing one launch on `sm_count` CTAs: equal tensor work per CTA.
void wgrad_multi_plan(int sm_count, int64_t M, int n, const int* N1, const int* N2, WgradPlan* out)
{
    double denom = 0.0;
    for (int j = 0; j < n; ++j) {
        out[j].nacc = (N1[j] % 256 == 0) ? 2 : 1;
        out[j].n1_blocks = N1[j] / (128 * out[j].nacc);
        denom += out[j].n1_blocks * out[j].nacc * mma_clk(N2[j]);
    }
    int used = 0;
    for (int j = 0; j < n; ++j) {
        int sp = (int)(sm_count * out[j].nacc * mma_clk(N2[j]) / denom);
        if (sp < 1) sp = 1;
        out[j].splits = sp;
        used += sp * out[j].n1_blocks;
    }
    // hand the leftover CTAs to the job whose CTAs currently carry the most work
    for (;;) {
        int best = -1;
        double worst = 0.0;
        for (int j = 0; j < n; ++j) {
            const double cost = (double)M / out[j].splits * out[j].nacc * mma_clk(N2[j]);
            if (used + out[j].n1_blocks <= sm_count && cost > worst) { worst = cost; best = j; }
        }
        if (best < 0) break;
        ++out[best].splits;
        used += out[best].n1_blocks;
    }
    for (int j = 0; j < n; ++j) {
        int64_t rps = (M + out[j].splits - 1) / out[j].splits;
        rps = (rps + WKC - 1) / WKC * WKC;
        out[j].rows_per_split = (int)rps;
        out[j].splits = (int)((M + rps - 1) / rps);
    }
}
}  // namespace

int dppo_tc2_wgrad_splits(dppo_ctx* ctx, int64_t M, int N1, int N2)
{
    WgradPlan p;
    wgrad_multi_plan(ctx->sm_count, M, 1, &N1, &N2, &p);
    return p.splits;
}

void dppo_tc2_wgrad_multi_splits(dppo_ctx* ctx, int64_t M, int n, const int* N1, const int* N2, int* splits_out)
{
    WgradPlan p[3];
    wgrad_multi_plan(ctx->sm_count, M, n, N1, N2, p);
    for (int j = 0; j < n; ++j) splits_out[j] = p[j].splits;
}

int dppo_tc2_wgrad_multi(dppo_ctx* ctx, int n, const float* const* Dm, const int* ldd, const float* const* Hm, const int* ldh,
                         float* const* partials, const int* splits, int64_t M, const int* N1, const int* N2, cudaStream_t st)
{
    if (n < 1 || n > 3) DPPO_FAIL(ctx, "tc2_wgrad: 1..3 jobs per launch");
    WgradPlan p[3];
    wgrad_multi_plan(ctx->sm_count, M, n, N1, N2, p);
    WgradJobs jobs;
    jobs.n = n;
    jobs.M = M;
    jobs.rn_hi = DPPO_DBG(ctx->tc_debug, 1024) ? 1 : 0;
    jobs.in_first = (ctx->row_sweep >> 4) & 1;
    jobs.prefetch = (ctx->tc_prefetch >> 2) & 1;
    int cta = 0, max_stage = 0;
    for (int j = 0; j < n; ++j) {
        if (!dppo_tc2_wgrad_supported(M, N1[j], N2[j])) DPPO_FAIL(ctx, "tc2_wgrad: unsupported shape M=%lld N1=%d N2=%d", (long long)M, N1[j], N2[j]);
        if (ldd[j] % 4 != 0 || ldh[j] % 4 != 0 || !al16(Dm[j]) || !al16(Hm[j]) || !al16(partials[j])) DPPO_FAIL(ctx, "tc2_wgrad: operands must be 16-byte aligned");
        if (p[j].splits != splits[j]) DPPO_FAIL(ctx, "tc2_wgrad: caller sized the partials of job %d for %d splits, plan has %d", j, splits[j], p[j].splits);
        if (p[j].nacc != p[0].nacc) DPPO_FAIL(ctx, "tc2_wgrad: jobs of one launch must agree on N1 %% 256");
        WgradJob& q = jobs.j[j];
        if (!dppo_make_tensor_map_2d(&q.tmD, Dm[j], M, N1[j], ldd[j], 32, WKC, 4) || !dppo_make_tensor_map_2d(&q.tmH, Hm[j], M, N2[j], ldh[j], 32, WKC, 4) ||
            !dppo_make_tensor_map_2d(&q.tmP, partials[j], (int64_t)p[j].splits * N1[j], N2[j], N2[j], 32, 32, 3))
            DPPO_FAIL(ctx, "tc2_wgrad: cuTensorMapEncodeTiled failed");
        q.N1 = N1[j]; q.N2 = N2[j]; q.n1_blocks = p[j].n1_blocks; q.rows_per_split = p[j].rows_per_split;
        q.cta_begin = cta; q.cta_count = p[j].n1_blocks * p[j].splits;
        cta += q.cta_count;
        const int sb = stage_bytes_of(p[j].nacc, N2[j]);
        if (sb > max_stage) max_stage = sb;
    }
    int stages = (200 * 1024) / max_stage;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    const size_t smem = (size_t)stages * max_stage + 1024;
    jobs.smem_bytes = (int)smem;
    if (p[0].nacc == 2) {
        cudaFuncSetAttribute(tc2_wgrad_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dppo_launch_pdl(ctx, tc2_wgrad_kernel<2>, dim3(cta), dim3(THREADS), smem, st, jobs);
    } else {
        cudaFuncSetAttribute(tc2_wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dppo_launch_pdl(ctx, tc2_wgrad_kernel<1>, dim3(cta), dim3(THREADS), smem, st, jobs);
    }
    DPPO_CHECK_LAUNCH(ctx, "tc2_wgrad_kernel");
    return 0;
}

int dppo_tc2_wgrad(dppo_ctx* ctx, const float* Dm, int ldd, const float* Hm, int ldh, float* partials, int splits, int64_t M, int N1,
                   int N2, cudaStream_t st)
{
    return dppo_tc2_wgrad_multi(ctx, 1, &Dm, &ldd, &Hm, &ldh, &partials, &splits, M, &N1, &N2, st);
}
