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
constexpr int BM = 128;                    // UMMA M
constexpr int A_HALF = BM * KC * 4;        // 8 KB: one image (hi or lo) of a 128-row x 16-k activation chunk
constexpr int G_STAGES = 3;
constexpr int MAXN = 512;
constexpr int STG_BLK = 32 * 128;          // one 32-row x 32-column fp32 staging block (SWIZZLE_128B layout)

// RH = 128-row halves per tile.
//   RH == 2: a CTA tile is 256 rows x n_tile columns; both halves consume every weight chunk, which halves the
//            L2 -> shared-memory weight traffic per MMA (at one half per CTA the 148 SMs ask the L2 for more than
//            it can deliver) and doubles the tensor work behind each ring stage.  Two TMEM accumulators, so the
//            epilogue of a tile is not overlapped with the next tile's MMAs (only with its loads and splits).
//   RH == 1: 128-row tiles, accumulators double-buffered (epilogue overlaps the next tile's MMAs).
// Dynamic shared memory: G_STAGES x [A_hi | A_lo | B_hi | B_lo], then one staging block per epilogue warp (two if RH == 1).
template <int EPI, int RH>
__global__ void __launch_bounds__(THREADS, 1)
tc2_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmC, const unsigned char* __restrict__ Wimg,
                const float* __restrict__ bias, const float* __restrict__ Hact, int ldh, float* __restrict__ colsum, int64_t M,
                int N, int K, int n_tile, int m_tiles, int dbg)
{
    extern __shared__ unsigned char dyn_raw[];
    __shared__ __align__(8) uint64_t full[G_STAGES], ready[G_STAGES], empty[G_STAGES], tfull[2], tempty[2];
    __shared__ uint32_t s_tmem;
    constexpr int A_IMG = RH * A_HALF;
    constexpr int NSET = RH == 1 ? 2 : 1;      // accumulator sets in TMEM
    constexpr int NBUF = RH == 1 ? 2 : 1;      // staging blocks per epilogue warp

    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~(uintptr_t)1023);
    const int b_img = n_tile * KC * 4;
    const int stage_bytes = 2 * A_IMG + 2 * b_img;
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
                const int next = tile + gridDim.x;
                const int next_m = next < total ? next / n_tiles : -1;
                for (int c = 0; c < chunks; ++c) {
                    // pull the next tile's activations into L2 while this tile computes
                    if (next_m >= 0 && next_m != m_blk) tma_prefetch_2d(&tmA, c * KC, next_m * RH * BM);
                    mbar_wait(&empty[s], ph ^ 1);
                    const bool skip_b = (dbg & 1) && (tile != (int)blockIdx.x || c >= G_STAGES);
                    mbar_expect_tx(&full[s], (uint32_t)(A_IMG + (skip_b ? 0 : 2 * b_img)));
                    unsigned char* st = dyn + s * stage_bytes;
                    tma_load_2d(st, &tmA, c * KC, m_blk * RH * BM, &full[s]);        // box: 16 k x (RH*128) rows
                    if (!skip_b) bulk_copy_g2s(st + 2 * A_IMG, wsrc + (int64_t)c * 2 * b_img, 2u * (uint32_t)b_img, &full[s]);
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
                const uint32_t set = it % NSET;
                mbar_wait(&tempty[set], ((it / NSET) & 1) ^ 1);        // epilogue has drained this accumulator set
                fence_after();
                for (int c = 0; c < chunks; ++c) {
                    mbar_wait(&full[s], ph);
                    mbar_wait(&ready[s], ph);
                    fence_after();
                    const uint32_t a_hi = smem_u32(dyn + s * stage_bytes), a_lo = a_hi + A_IMG;
                    const uint32_t b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + (uint32_t)b_img;
#pragma unroll
                    for (int ks = 0; ks < KC / 8; ++ks) {
                        const uint32_t ko = ks * 32;                    // 8 tf32 = 32 bytes along K inside the swizzle atom
                        const uint64_t dbh = desc_k_sw64(b_hi + ko), dbl = desc_k_sw64(b_lo + ko);
#pragma unroll
                        for (int h = 0; h < RH; ++h) {
                            const uint32_t d = tmem + (set * RH + h) * 256;
                            const uint64_t dah = desc_k_sw64(a_hi + h * A_HALF + ko), dal = desc_k_sw64(a_lo + h * A_HALF + ko);
                            if (!(dbg & 4)) {
                                umma_tf32(d, dah, dbl, idesc, (c | ks) != 0);
                                umma_tf32(d, dal, dbh, idesc, 1u);
                                umma_tf32(d, dah, dbh, idesc, 1u);
                            } else {
                                umma_tf32(d, dah, dbh, idesc, (c | ks) != 0);
                            }
                        }
                    }
                    umma_commit(&empty[s]);                             // frees the stage once these MMAs have read it
                    if (++s == G_STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[set]);
            }
        }
        __syncwarp();
    } else if (warp < W_EPI0) {
        // hi/lo split of the activation chunk TMA delivered (raw fp32 lands in the hi image)
        const int ct = tid - W_SPLIT0 * 32;
        constexpr int PER = A_IMG / 16 / SPLIT_THREADS;
        int s = 0;
        uint32_t ph = 0;
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
            for (int c = 0; c < chunks; ++c) {
                mbar_wait(&full[s], ph);
                const uint32_t hi = smem_u32(dyn + s * stage_bytes) + ct * 16;
                float4 x[PER];
#pragma unroll
                for (int j = 0; j < PER; ++j) x[j] = lds128(hi + j * SPLIT_THREADS * 16);
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const float4 l = split_tf32x4(x[j]);
                    sts128(hi + j * SPLIT_THREADS * 16, x[j]);
                    sts128(hi + A_IMG + j * SPLIT_THREADS * 16, l);
                }
                fence_proxy_async();                                    // generic-proxy writes -> visible to the tensor core
                mbar_arrive(&ready[s]);
                if (++s == G_STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else {
        // epilogue: warp -> TMEM lane quadrant (warp % 4); the two warps of a quadrant split the tile by row half (RH == 2)
        // or by column half (RH == 1).  Results leave through a per-warp 32x32 staging block in the SWIZZLE_128B layout and
        // a TMA store, so that HBM sees full 128-byte row segments (a thread-per-row st.global touches 32 different lines
        // per instruction and was measured to be the bottleneck).
        const int ew = warp - W_EPI0;
        const int q = warp & 3, grp = ew >> 2;
        const int h = RH == 2 ? grp : 0;
        const int ncols = RH == 2 ? n_tile : n_tile / 2;
        const int col0 = RH == 2 ? 0 : grp * ncols;
        const int nblk = ncols / 32;
        const uint32_t stg = smem_u32(dyn + G_STAGES * stage_bytes) + (uint32_t)ew * NBUF * STG_BLK;
        const uint32_t row_off = (uint32_t)lane * 128, sw = (uint32_t)(lane & 7);
        uint32_t it = 0, blk = 0;
        if (lane == 0) tma_prefetch_desc(&tmC);
        for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
            const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
            const int n0 = n_blk * n_tile + col0;
            const int m0 = (m_blk * RH + h) * BM + q * 32;
            const int64_t m = (int64_t)m0 + lane;
            const uint32_t set = it % NSET;
            mbar_wait(&tfull[set], (it / NSET) & 1);
            fence_after();
            for (int k = 0; k < ((dbg & 8) ? 0 : nblk); ++k, ++blk) {
                const int n = n0 + k * 32;
                const uint32_t buf = stg + (blk % NBUF) * STG_BLK;
                uint32_t r[32];
                tmem_ld32_issue(tmem + (set * RH + h) * 256 + ((uint32_t)(q * 32) << 16) + (uint32_t)(col0 + k * 32), r);
                float4 aux[8];                                          // bias (forward) or the layer's activations (dgrad)
                if (EPI == DPPO_EPI_BIAS_TANH) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) aux[j] = __ldg(reinterpret_cast<const float4*>(bias + n) + j);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        aux[j] = m < M ? __ldg(reinterpret_cast<const float4*>(Hact + m * (int64_t)ldh + n) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                tmem_ld32_wait(r);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (EPI == DPPO_EPI_BIAS_TANH) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (dbg & 32) {
                            v[4 * j] += aux[j].x; v[4 * j + 1] += aux[j].y; v[4 * j + 2] += aux[j].z; v[4 * j + 3] += aux[j].w;
                        } else {
                            v[4 * j] = tanhf(v[4 * j] + aux[j].x); v[4 * j + 1] = tanhf(v[4 * j + 1] + aux[j].y);
                            v[4 * j + 2] = tanhf(v[4 * j + 2] + aux[j].z); v[4 * j + 3] = tanhf(v[4 * j + 3] + aux[j].w);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[4 * j] *= (1.0f - aux[j].x * aux[j].x); v[4 * j + 1] *= (1.0f - aux[j].y * aux[j].y);
                        v[4 * j + 2] *= (1.0f - aux[j].z * aux[j].z); v[4 * j + 3] *= (1.0f - aux[j].w * aux[j].w);
                    }
                }
                // the store that last read this staging block must be done with it
                if (lane == 0) { if (NBUF == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    sts128(buf + row_off + ((((uint32_t)j) ^ sw) << 4), make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0 && !(dbg & 16)) {
                    tma_store_2d(&tmC, n, m0, buf);                     // rows >= M are clipped by the tensor map
                    bulk_commit();
                }
                if (EPI == DPPO_EPI_TANH_BWD && colsum != nullptr) {
                    // rows >= M hold exact zeros (TMA zero-fills the out-of-range A rows), so they do not disturb the sums.
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
                    colsum[((int64_t)(m_blk * RH + h) * 4 + q) * N + n + lane] = v[0];
                }
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[set]);
        }
        if (lane == 0) bulk_wait<0>();             // staging memory must outlive the last store's read
        __syncwarp();
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
            for (int c = 0; c < PF && c < chunks; ++c) {
                const int row = (int)(r0 + (int64_t)c * WKC);
                for (int g = 0; g < gA; ++g) tma_prefetch_2d(&tmD, n1_0 + g * 32, row);
                for (int g = 0; g < gB; ++g) tma_prefetch_2d(&tmH, g * 32, row);
            }
            for (int c = 0; c < chunks; ++c) {
                if (c + PF < chunks) {
                    const int prow = (int)(r0 + (int64_t)(c + PF) * WKC);
                    for (int g = 0; g < gA; ++g) tma_prefetch_2d(&tmD, n1_0 + g * 32, prow);
                    for (int g = 0; g < gB; ++g) tma_prefetch_2d(&tmH, g * 32, prow);
                }
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

// ---- launchers -------------------------------------------------------------------------------------
bool dppo_tc2_gemm_supported(int64_t M, int N, int K)
{
    return M >= 1024 && M < (int64_t)1 << 31 && K % KC == 0 && (N % 256 == 0 || N == 128) && N <= MAXN;
}

int dppo_tc2_colsum_parts(int64_t M) { return (int)((M + 2 * BM - 1) / (2 * BM)) * 8; }

int dppo_tc2_gemm(dppo_ctx* ctx, int epi, const float* A, int lda, const unsigned char* Wimg, const float* bias, const float* Hact,
                  int ldh, float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st)
{
    if (!dppo_tc2_gemm_supported(M, N, K)) DPPO_FAIL(ctx, "tc2_gemm: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    if (lda % 4 != 0 || ldc % 4 != 0 || !al16(A) || !al16(C) || !al16(Wimg) || (Hact && (!al16(Hact) || ldh % 4 != 0)) || (bias && !al16(bias)))
        DPPO_FAIL(ctx, "tc2_gemm: operands must be 16-byte aligned with row pitches multiple of 4 floats");
    const int rh = (ctx->tc_debug & 64) ? 1 : 2;
    CUtensorMap tmA, tmC;
    if (!dppo_make_tensor_map_2d(&tmA, A, M, K, lda, KC, rh * BM, 2) || !dppo_make_tensor_map_2d(&tmC, C, M, N, ldc, 32, 32, 3))
        DPPO_FAIL(ctx, "tc2_gemm: cuTensorMapEncodeTiled failed");
    const int n_tile = dppo_tc_n_tile(N);
    const int m_tiles = (int)((M + rh * BM - 1) / (rh * BM));
    const int total = m_tiles * (N / n_tile);
    const size_t smem = (size_t)G_STAGES * (2 * rh * A_HALF + 2 * n_tile * KC * 4) + (size_t)N_EPI * (rh == 1 ? 2 : 1) * STG_BLK + 1024;
    const int grid = total < ctx->sm_count ? total : ctx->sm_count;
#define TC2_LAUNCH(EPI, RH)                                                                                                   \
    do {                                                                                                                      \
        cudaFuncSetAttribute(tc2_gemm_kernel<EPI, RH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
        tc2_gemm_kernel<EPI, RH><<<grid, THREADS, smem, st>>>(tmA, tmC, Wimg, bias, Hact, ldh, colsum, M, N, K, n_tile, m_tiles, \
                                                              ctx->tc_debug);                                                 \
    } while (0)
    if (epi == DPPO_EPI_BIAS_TANH) {
        if (rh == 2) TC2_LAUNCH(DPPO_EPI_BIAS_TANH, 2); else TC2_LAUNCH(DPPO_EPI_BIAS_TANH, 1);
    } else if (epi == DPPO_EPI_TANH_BWD) {
        if (rh == 2) TC2_LAUNCH(DPPO_EPI_TANH_BWD, 2); else TC2_LAUNCH(DPPO_EPI_TANH_BWD, 1);
    } else {
        DPPO_FAIL(ctx, "tc2_gemm: unknown epilogue %d", epi);
    }
#undef TC2_LAUNCH
    DPPO_CHECK_LAUNCH(ctx, "tc2_gemm_kernel");
    return 0;
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
    jobs.rn_hi = (ctx->tc_debug & 1024) ? 1 : 0;
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
