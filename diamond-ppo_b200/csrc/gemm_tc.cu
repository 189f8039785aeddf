// Tensor-core GEMMs for the wide (H >= 128) actor-critic MLP: 3xTF32 on tcgen05 with TMEM accumulators.
//
// fp32 parity needs more than one TF32 pass (single-pass TF32 misses the 1e-4 parameter bar by 70x, SURVEY.md
// §0.6), so every fp32 operand x is split into hi = x rounded to the nearest TF32 number (10 mantissa bits)
// and lo = x - hi (exact in fp32; the tensor core reads its top 19 bits), and each k-step issues
// three MMAs into the same fp32 TMEM accumulator:   D += A_hi*B_lo ; D += A_lo*B_hi ; D += A_hi*B_hi.
// The dropped A_lo*B_lo term and the truncation of lo are O(2^-22) relative per product.
//
//   C[M, N] = epi( A[M, K] * Wimg )     A: fp32 row-major activations (optionally row-gathered),
//                                       Wimg: weights pre-split and pre-swizzled by prep_weights_kernel
// One CTA = one 128-row x N_TILE(<=256) output tile; UMMA M=128, N=N_TILE, K=8 (kind::tf32), cta_group::1;
// two CTAs per SM share the tensor core, so one CTA's epilogue overlaps the other's main loop.
// Operands live in shared memory in the canonical K-major SWIZZLE_64B layout (16 fp32 of K per 64-byte row,
// 8-row atoms of 512 B).  Per 16-wide k-chunk: all threads split their part of the A chunk (prefetched into
// registers one chunk ahead) into the hi/lo images; one thread streams the 2 x N_TILE x 64 B weight images
// with a bulk async copy (TMA engine, mbarrier complete_tx); one thread issues 6 MMAs and commits them to the
// stage's mbarrier, which is what frees the stage for re-use two chunks later.
#include <cuda.h>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace {

constexpr int TC_THREADS = 256;
constexpr int BM = 128;                    // rows per CTA tile (UMMA M)
constexpr int KC = 16;                     // k elements per chunk (one SWIZZLE_64B atom width)
constexpr int A_IMG = BM * KC * 4;         // 8 KB: one A image (hi or lo) of a chunk
constexpr int STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TC_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TC_DONE_%=;\n\t"
        "bra TC_WAIT_%=;\n\t"
        "TC_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major SWIZZLE_64B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 512 B between 8-row groups, version 1 (sm_100), layout 4 = SWIZZLE_64B.
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t saddr)
{
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (4ull << 61);
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, both K-major, N>>3, M>>4.
__device__ __forceinline__ uint32_t make_idesc(int n)
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32])
{
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// byte offset of element (row, kk) inside a K-major SWIZZLE_64B image (rows x 16 fp32)
__host__ __device__ __forceinline__ int sw64_offset(int row, int kk)
{
    return (row >> 3) * 512 + (row & 7) * 64 + ((((kk >> 2) ^ ((row & 7) >> 1)) & 3) << 4) + ((kk & 3) << 2);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo)
{
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);   // round to nearest TF32 (|lo| <= 2^-12 |x|)
    lo = x - hi;
}

// ---------------------------------------------------------------------------------------------
// Weight images.  For one Linear weight W [rows_w, cols_w] (row-major) used as the B operand
//   transpose == 0 : B[n, k] = W[n, k]      (forward:  N = rows_w, K = cols_w)
//   transpose == 1 : B[n, k] = W[k, n]      (dgrad:    N = cols_w, K = rows_w)
// Output: for every N tile (n_tile rows) and every 16-wide k chunk, [hi image | lo image], each
// n_tile x 64 B in the SWIZZLE_64B layout, so that the GEMM copies a chunk with one bulk copy.
// ---------------------------------------------------------------------------------------------
__global__ void prep_weights_kernel(const float* __restrict__ W, int rows_w, int cols_w, int transpose, int n_tile,
                                    unsigned char* __restrict__ img)
{
    const int N = transpose ? cols_w : rows_w;
    const int K = transpose ? rows_w : cols_w;
    const int64_t total = (int64_t)N * K;
    const int chunks = K / KC;
    const int img_bytes = n_tile * KC * 4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int n, k;
        if (transpose) { k = (int)(e / cols_w); n = (int)(e % cols_w); }      // coalesced read of W[k, n]
        else { n = (int)(e / cols_w); k = (int)(e % cols_w); }
        const float x = W[e];
        float hi, lo;
        split_tf32(x, hi, lo);
        const int tile = n / n_tile, row = n % n_tile, chunk = k / KC, kk = k % KC;
        unsigned char* base = img + ((int64_t)(tile * chunks + chunk) * 2) * img_bytes;
        const int off = sw64_offset(row, kk);
        *reinterpret_cast<float*>(base + off) = hi;
        *reinterpret_cast<float*>(base + img_bytes + off) = lo;
    }
}

// All weight images of one optimiser step in a single launch: blockIdx.y selects the job.
__global__ void prep_weights_multi_kernel(PrepJobs jobs)
{
    DPPO_PDL_ENTER();
    if ((int)blockIdx.y >= jobs.n) {
        // the minibatch observation gather (ppo.py:261 observations[mb]) shares the launch: it is independent of the weights.
        // It takes the remaining gridDim.y - jobs.n slices of the grid (index -> row is a dependent pair of DRAM latencies, so the
        // work is spread over more threads instead of more iterations per thread); two items in flight per thread
        const GatherJob& g = jobs.gather;
        const int64_t total = g.rows * g.row_vec;
        const int64_t nthreads = (int64_t)(gridDim.y - jobs.n) * gridDim.x * blockDim.x;
        const int64_t t0 = ((int64_t)(blockIdx.y - jobs.n) * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
        for (int64_t i = t0; i < total; i += 2 * nthreads) {
            const int64_t i2 = i + nthreads;
            const int64_t r = i / g.row_vec, r2 = i2 / g.row_vec;
            const int32_t s1 = __ldg(g.idx + r), s2 = i2 < total ? __ldg(g.idx + r2) : 0;
            const float4 v1 = __ldg(g.src + (int64_t)s1 * g.row_vec + (int)(i - r * g.row_vec));
            if (i2 < total) {
                const float4 v2 = __ldg(g.src + (int64_t)s2 * g.row_vec + (int)(i2 - r2 * g.row_vec));
                g.dst[i2] = v2;
            }
            g.dst[i] = v1;
        }
        return;
    }
    const PrepJob& j = jobs.job[blockIdx.y];
    const int N = j.transpose ? j.cols_w : j.rows_w;
    const int K = j.transpose ? j.rows_w : j.cols_w;
    const int64_t total = (int64_t)N * K;
    const int chunks = K / KC;
    const int img_bytes = j.n_tile * KC * 4;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int n, k;
        if (j.transpose) { k = (int)(e / j.cols_w); n = (int)(e % j.cols_w); }
        else { n = (int)(e / j.cols_w); k = (int)(e % j.cols_w); }
        const float x = j.W[e];
        float hi, lo;
        split_tf32(x, hi, lo);
        const int tile = n / j.n_tile, row = n % j.n_tile, chunk = k / KC, kk = k % KC;
        unsigned char* base = j.img + ((int64_t)(tile * chunks + chunk) * 2) * img_bytes;
        const int off = sw64_offset(row, kk);
        *reinterpret_cast<float*>(base + off) = hi;
        *reinterpret_cast<float*>(base + img_bytes + off) = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// The GEMM.  EPI: DPPO_EPI_BIAS_TANH (forward) or DPPO_EPI_TANH_BWD (dgrad, + column sums).
// grid = (m_tiles, n_tiles).  Dynamic smem: STAGES * (2*A_IMG + 2*n_tile*64) + 1024 alignment slack.
// ---------------------------------------------------------------------------------------------
template <int EPI>
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_gemm_kernel(const float* __restrict__ A, int lda, const int32_t* __restrict__ a_rows, const unsigned char* __restrict__ Wimg,
               const float* __restrict__ bias, const float* __restrict__ Hact, int ldh, float* __restrict__ C, int ldc,
               float* __restrict__ colsum, int64_t M, int N, int K, int n_tile)
{
    extern __shared__ unsigned char dyn_raw[];
    __shared__ __align__(8) uint64_t b_full[STAGES];
    __shared__ __align__(8) uint64_t mma_done[STAGES];
    __shared__ uint32_t s_tmem;
    __shared__ float s_col[4][256];

    unsigned char* dyn = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~(uintptr_t)1023);
    const int b_img = n_tile * KC * 4;                   // bytes of one B image
    const int stage_bytes = 2 * A_IMG + 2 * b_img;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n_blk = blockIdx.y;
    const int n0 = n_blk * n_tile;
    const int chunks = K / KC;
    const unsigned char* wsrc = Wimg + (int64_t)n_blk * chunks * 2 * b_img;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&mma_done[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    // A staging: thread -> (row = tid/2, 8 consecutive k = (tid&1)*8 ..)
    const int a_row = tid >> 1, a_k0 = (tid & 1) * 8;
    const int64_t gm = m0 + a_row;
    const bool a_ok = gm < M;
    const float* a_ptr = A + (a_ok ? (a_rows ? (int64_t)a_rows[gm] : gm) : 0) * (int64_t)lda + a_k0;
    float4 ra0 = make_float4(0.f, 0.f, 0.f, 0.f), ra1 = ra0;
    auto load_a = [&](int c) {
        if (a_ok) {
            ra0 = __ldg(reinterpret_cast<const float4*>(a_ptr + c * KC));
            ra1 = __ldg(reinterpret_cast<const float4*>(a_ptr + c * KC + 4));
        }
    };
    const int a_off0 = sw64_offset(a_row, a_k0), a_off1 = sw64_offset(a_row, a_k0 + 4);
    auto store_a = [&](int s) {
        unsigned char* hi = dyn + s * stage_bytes;
        unsigned char* lo = hi + A_IMG;
        float4 h, l;
        split_tf32(ra0.x, h.x, l.x); split_tf32(ra0.y, h.y, l.y); split_tf32(ra0.z, h.z, l.z); split_tf32(ra0.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi + a_off0) = h;
        *reinterpret_cast<float4*>(lo + a_off0) = l;
        split_tf32(ra1.x, h.x, l.x); split_tf32(ra1.y, h.y, l.y); split_tf32(ra1.z, h.z, l.z); split_tf32(ra1.w, h.w, l.w);
        *reinterpret_cast<float4*>(hi + a_off1) = h;
        *reinterpret_cast<float4*>(lo + a_off1) = l;
    };
    auto issue_b = [&](int c, int s) {
        mbar_expect_tx(&b_full[s], 2u * (uint32_t)b_img);
        bulk_copy_g2s(dyn + s * stage_bytes + 2 * A_IMG, wsrc + (int64_t)c * 2 * b_img, 2u * (uint32_t)b_img, &b_full[s]);
    };

    if (tid == 0) {
        issue_b(0, 0);
        if (chunks > 1) issue_b(1, 1);
    }
    load_a(0);
    const uint32_t idesc = make_idesc(n_tile);

    for (int c = 0; c < chunks; ++c) {
        const int s = c & 1;
        const uint32_t use = (uint32_t)(c >> 1);                   // how many times this stage was used before
        if (c >= STAGES) {
            mbar_wait(&mma_done[s], (use - 1) & 1);                // MMAs that read this stage have completed
            if (tid == 0) issue_b(c, s);
        }
        store_a(s);
        if (c + 1 < chunks) load_a(c + 1);                         // prefetch overlaps the MMAs issued below
        fence_proxy_async();                                       // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&b_full[s], use & 1);
            tc_fence_after();
            const uint32_t a_hi = smem_u32(dyn + s * stage_bytes), a_lo = a_hi + A_IMG;
            const uint32_t b_hi = a_hi + 2 * A_IMG, b_lo = b_hi + (uint32_t)b_img;
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
                const uint32_t ko = ks * 32;                        // 8 tf32 = 32 bytes along K inside the swizzle atom
                umma_tf32(tmem, make_desc_sw64(a_hi + ko), make_desc_sw64(b_lo + ko), idesc, (c | ks) != 0);
                umma_tf32(tmem, make_desc_sw64(a_lo + ko), make_desc_sw64(b_hi + ko), idesc, 1u);
                umma_tf32(tmem, make_desc_sw64(a_hi + ko), make_desc_sw64(b_hi + ko), idesc, 1u);
            }
            umma_commit(&mma_done[s]);
        }
    }
    // accumulator complete when the last chunk's MMAs (issued in order) have committed
    {
        const int last = chunks - 1;
        mbar_wait(&mma_done[last & 1], (uint32_t)(last >> 1) & 1);
        tc_fence_after();
    }

    // ---- epilogue: warp -> TMEM lane quadrant (warp%4) and column half (warp/4) ----
    const int q = warp & 3, half = warp >> 2;
    const int64_t m = m0 + q * 32 + lane;
    const int cols_per_half = n_tile / 2;
    for (int cb = 0; cb < cols_per_half; cb += 32) {
        const int col = half * cols_per_half + cb;
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)col, v);
        const int n = n0 + col;
        if (EPI == DPPO_EPI_BIAS_TANH) {
            if (m < M) {
                float* dst = C + m * (int64_t)ldc + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    o.x = tanhf(v[j] + __ldg(bias + n + j));
                    o.y = tanhf(v[j + 1] + __ldg(bias + n + j + 1));
                    o.z = tanhf(v[j + 2] + __ldg(bias + n + j + 2));
                    o.w = tanhf(v[j + 3] + __ldg(bias + n + j + 3));
                    *reinterpret_cast<float4*>(dst + j) = o;
                }
            }
        } else {
            if (m < M) {
                const float* hp = Hact + m * (int64_t)ldh + n;
                float* dst = C + m * (int64_t)ldc + n;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 h = __ldg(reinterpret_cast<const float4*>(hp + j));
                    v[j] *= (1.0f - h.x * h.x); v[j + 1] *= (1.0f - h.y * h.y);
                    v[j + 2] *= (1.0f - h.z * h.z); v[j + 3] *= (1.0f - h.w * h.w);
                    *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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
                s_col[q][col + lane] = v[0];
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (EPI == DPPO_EPI_TANH_BWD && colsum != nullptr) {
        for (int j = tid; j < n_tile; j += TC_THREADS)
            colsum[(int64_t)blockIdx.x * N + n0 + j] = (s_col[0][j] + s_col[1][j]) + (s_col[2][j] + s_col[3][j]);
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

}  // namespace

bool dppo_tc_supported(int64_t M, int N, int K)
{
    return M >= 1024 && K % KC == 0 && (N % 256 == 0 || N == 128);
}

int dppo_tc_n_tile(int N) { return N % 256 == 0 ? 256 : 128; }

int64_t dppo_tc_image_bytes(int N, int K) { return (int64_t)N * K * 4 * 2; }

int dppo_tc_prep_weights(dppo_ctx* ctx, const float* W, int rows_w, int cols_w, int transpose, unsigned char* img, cudaStream_t st)
{
    const int N = transpose ? cols_w : rows_w;
    const int64_t total = (int64_t)rows_w * cols_w;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
    prep_weights_kernel<<<blocks, 256, 0, st>>>(W, rows_w, cols_w, transpose, dppo_tc_n_tile(N), img);
    DPPO_CHECK_LAUNCH(ctx, "prep_weights_kernel");
    return 0;
}

int dppo_tc_prep_weights_multi(dppo_ctx* ctx, PrepJobs jobs, cudaStream_t st)
{
    const bool with_gather = jobs.gather.rows > 0 && jobs.gather.dst != nullptr;
    if (jobs.n < 1 && !with_gather) return 0;
    int64_t most = with_gather ? jobs.gather.rows * jobs.gather.row_vec / 8 : 0;
    for (int i = 0; i < jobs.n; ++i) {
        const int64_t t = (int64_t)jobs.job[i].rows_w * jobs.job[i].cols_w;
        jobs.job[i].n_tile = dppo_tc_n_tile(jobs.job[i].transpose ? jobs.job[i].cols_w : jobs.job[i].rows_w);
        if (t > most) most = t;
    }
    int blocks = (int)((most + 255) / 256);
    if (blocks > 4 * ctx->sm_count) blocks = 4 * ctx->sm_count;
    if (blocks < 1) blocks = 1;
    dppo_launch_pdl(ctx, prep_weights_multi_kernel, dim3(blocks, jobs.n + (with_gather ? 4 : 0)), dim3(256), 0, st, jobs);
    DPPO_CHECK_LAUNCH(ctx, "prep_weights_multi_kernel");
    return 0;
}

int dppo_tc_gemm(dppo_ctx* ctx, int epi, const float* A, int lda, const int32_t* a_rows, const unsigned char* Wimg, const float* bias,
                 const float* Hact, int ldh, float* C, int ldc, float* colsum, int64_t M, int N, int K, cudaStream_t st)
{
    if (!dppo_tc_supported(M, N, K)) DPPO_FAIL(ctx, "tc_gemm: unsupported shape M=%lld N=%d K=%d", (long long)M, N, K);
    if (lda % 4 != 0 || ldc % 4 != 0 || (reinterpret_cast<uintptr_t>(A) & 15u) || (reinterpret_cast<uintptr_t>(C) & 15u) ||
        (reinterpret_cast<uintptr_t>(Wimg) & 15u))
        DPPO_FAIL(ctx, "tc_gemm: operands must be 16-byte aligned with row pitches multiple of 4 floats");
    const int n_tile = dppo_tc_n_tile(N);
    const size_t smem = (size_t)STAGES * (2 * A_IMG + 2 * n_tile * KC * 4) + 1024;
    dim3 grid((unsigned)((M + BM - 1) / BM), (unsigned)(N / n_tile));
    if (epi == DPPO_EPI_BIAS_TANH) {
        cudaFuncSetAttribute(tc_gemm_kernel<DPPO_EPI_BIAS_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_kernel<DPPO_EPI_BIAS_TANH><<<grid, TC_THREADS, smem, st>>>(A, lda, a_rows, Wimg, bias, nullptr, 0, C, ldc, nullptr, M, N, K, n_tile);
    } else {
        cudaFuncSetAttribute(tc_gemm_kernel<DPPO_EPI_TANH_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tc_gemm_kernel<DPPO_EPI_TANH_BWD><<<grid, TC_THREADS, smem, st>>>(A, lda, nullptr, Wimg, nullptr, Hact, ldh, C, ldc, colsum, M, N, K, n_tile);
    }
    DPPO_CHECK_LAUNCH(ctx, "tc_gemm_kernel");
    return 0;
}
