// Dense tall-skinny FP64 GEMM on the tensor cores:
//   trans = 0:  W[mloc x lp] = alpha * A[mloc x n]  * X[n x lp]      (`A * Omega`, `A * Q`)
//   trans = 1:  W[n x lp]    = alpha * A[mloc x n]' * X[mloc x lp]   (`A' * Q`, `(Q' * A)'`)
// A is column-major in HBM (Julia layout) and is read exactly once per pass through
// tiled TMA (cp.async.bulk.tensor.2d -> UTMALDG) into shared memory; the box is
// over-fetched by 4 elements along its inner dimension so that the tile pitch is
// 4 (mod 16) doubles and the DMMA A-fragment LDS.64 pattern is bank-conflict free
// without swizzling.  X tiles (TALL layout, contiguous rows) arrive by 1-D bulk copy.
// Math: mma.sync.m8n8k4.f64 (DMMA.8x8x4); 16 warps per CTA, each accumulating a
// 16 x (8*NB/4) strip in registers (one warp issues a DMMA only every ~32 cycles, so the SM
// needs 3-4 warps per scheduler to keep the FP64 tensor pipe busy -- profiles/r01).
#include "common.cuh"
#include "algos.h"
#include "ptx.cuh"
#include "nb_list.h"

namespace gsi {

constexpr int DG_BM = 64;
constexpr int DG_BK = 32;
constexpr int DG_PAD = 4;
constexpr int DG_CONSUMERS = 16;                // 4 row groups (16 rows) x 4 column groups, as in kcov_gemm
constexpr int DG_THREADS = DG_CONSUMERS * 32;   // thread 0 doubles as the TMA producer
constexpr int DG_CG = 4;

struct DenseParams {
    const double* X;
    double* W;
    int64_t out_rows;     // rows of W (mloc for N, n for T)
    int64_t kdim;         // reduction length (n for N, mloc for T)
    int64_t ld, ldw;
    double alpha;
    int stages;
    // split-K (short-and-wide products: few 64-row output tiles, long reduction -- S'X of the LowRankCovMatrix,
    // the rga sketch S*V, Q'Y): CTA b works on output tile b / ksplit and k-tiles [part * kt_per, ...) with
    // part = b % ksplit, and stores raw partial sums to partial[part]; dense_splitk_fixup_kernel adds the parts
    // in ascending order.  ksplit == 1: the ordinary schedule below.
    int ksplit;
    int64_t kt_per;
    double* partial;      // [ksplit][out_rows][ldw]
};

template <int NB, int TRANS>
__global__ void __launch_bounds__(DG_THREADS, 1)
dense_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ DenseParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ld = NB * 8 + 4;
    constexpr int a_inner = (TRANS ? DG_BK : DG_BM) + DG_PAD;      // pitch of the A tile (doubles)
    constexpr int a_outer = TRANS ? DG_BM : DG_BK;
    constexpr int a_doubles = a_inner * a_outer;
    constexpr int stage_doubles = DG_BK * ld + a_doubles;
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_doubles);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nstages = p.stages;
    if (tid == 0) {
        tma_prefetch_desc(&amap);
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], DG_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Work schedule (the one of kcov_gemm.cu): full rounds of 64-row tiles, CTA b taking tile (round * grid + b),
    // then ONE tail round that spreads the remaining (< 4 * grid) 16-row groups over all CTAs, q = ceil(remaining /
    // grid) groups each -- a CTA with fewer active row groups finishes its k sweep sooner (its warps are issue-bound,
    // not pipe-bound), which removes most of the wave-quantisation loss (n = 32768: 512 tiles on 148 SMs ran as 4
    // rounds at 86 % efficiency; ncu of that launch: DMMA pipe 95 % busy, 28.7 TF/s).
    const int64_t total_rg = (p.out_rows + 15) / 16;
    const bool splitk = p.ksplit > 1;
    const int64_t nslots = splitk ? gridDim.x / p.ksplit : gridDim.x;        // CTAs (or CTA groups) sharing the rows
    const int64_t slot = splitk ? blockIdx.x / p.ksplit : blockIdx.x;
    const int part = splitk ? (int)(blockIdx.x % p.ksplit) : 0;
    const int64_t per_round = nslots * 4;
    const int64_t full_rounds = splitk ? 0 : total_rg / per_round;
    const int64_t remaining = total_rg - full_rounds * per_round;
    const int64_t q_tail = splitk ? 4 : (remaining + nslots - 1) / nslots;   // 0..4 (split-K: whole 64-row tiles)
    const bool has_tail = remaining > 0 && slot * q_tail < remaining;
    const int64_t my_rounds = full_rounds + (has_tail ? 1 : 0);
    auto round_base = [&](int64_t j, int& nact) -> int64_t {                 // first row group, active row groups
        if (j < full_rounds) { nact = 4; return (j * nslots + slot) * 4; }
        const int64_t b0 = slot * q_tail;
        const int64_t left = remaining - b0;
        nact = (int)(left < q_tail ? left : q_tail);
        return full_rounds * per_round + b0;
    };
    const int64_t nkt_all = (p.kdim + DG_BK - 1) / DG_BK;
    const int64_t kt_lo = splitk ? part * p.kt_per : 0;                      // my k-tiles [kt_lo, kt_lo + nkt)
    const int64_t nkt = splitk ? ((nkt_all - kt_lo < p.kt_per) ? nkt_all - kt_lo : p.kt_per) : nkt_all;
    constexpr uint32_t stage_bytes = (uint32_t)(stage_doubles * sizeof(double));

    const int64_t total_it = my_rounds * nkt;
    const int lookahead = nstages > 2 ? nstages - 2 : 1;
    // producer state, advanced incrementally (no 64-bit divisions on thread 0's path -- its warp is an MMA warp too)
    int64_t p_it = 0, p_kt = 0, p_round = 0, p_row = 0;
    int p_s = 0;
    uint32_t p_ph = 0;
    {
        int nact_unused;
        if (my_rounds > 0) p_row = round_base(0, nact_unused) * 16;       // first output row of the tile
    }
    auto produce_next = [&]() {
        mbar_wait(&empty[p_s], p_ph ^ 1u);
        double* xs = smem + (size_t)p_s * stage_doubles;
        double* as = xs + DG_BK * ld;
        mbar_expect_tx(&full[p_s], stage_bytes);
        const int64_t ktg = kt_lo + p_kt;                                     // global k-tile
        bulk_g2s(xs, p.X + ktg * DG_BK * p.ld, DG_BK * ld * 8, &full[p_s]);
        if (TRANS) tma_load_2d(as, &amap, (int)(ktg * DG_BK), (int)p_row, &full[p_s]);
        else       tma_load_2d(as, &amap, (int)p_row, (int)(ktg * DG_BK), &full[p_s]);
        ++p_it;
        if (++p_kt == nkt) {
            p_kt = 0;
            ++p_round;
            int nact_unused;
            if (p_round < my_rounds) p_row = round_base(p_round, nact_unused) * 16;
        }
        if (++p_s == nstages) { p_s = 0; p_ph ^= 1u; }
    };
    if (tid == 0) {
        for (int64_t i = 0; i < lookahead && i < total_it; ++i) produce_next();
    }

    const int g = lane >> 2, t = lane & 3;
    constexpr int NBW = (NB + DG_CG - 1) / DG_CG;
    const int rg = warp >> 2;
    const int cg = (warp + rg) & 3;
    const int nb0 = cg * NBW;
    int c_s = 0;                       // consumer pipeline stage / mbarrier phase
    uint32_t c_ph = 0;
    for (int64_t round = 0; round < my_rounds; ++round) {
        int nact = 0;
        const int64_t base_rg = round_base(round, nact);
        const bool active = rg < nact;                                       // warp-uniform
        double acc[2][NBW][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0.0; acc[h][nb][1] = 0.0; }
        for (int64_t kt = 0; kt < nkt; ++kt) {
            if (tid == 0 && p_it < total_it) produce_next();
            const int s = c_s;
            mbar_wait(&full[s], c_ph);
            if (++c_s == nstages) { c_s = 0; c_ph ^= 1u; }
            __syncwarp();
            const double* xs = smem + (size_t)s * stage_doubles;
            const double* as = xs + DG_BK * ld;
            if (active) {
#pragma unroll
                for (int ks = 0; ks < DG_BK / 4; ++ks) {
                    const int j = ks * 4 + t;
                    const int r = rg * 16 + g;
                    const double a0 = TRANS ? as[r * a_inner + j] : as[j * a_inner + r];
                    const double a1 = TRANS ? as[(r + 8) * a_inner + j] : as[j * a_inner + r + 8];
                    const double* xrow = xs + j * ld + nb0 * 8 + g;
#pragma unroll
                    for (int nb = 0; nb < NBW; ++nb) {
                        if (nb0 + nb < NB) {
                            const double b = xrow[nb * 8];
                            dmma884(acc[0][nb][0], acc[0][nb][1], a0, b);
                            dmma884(acc[1][nb][0], acc[1][nb][1], a1, b);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t row = (base_rg + rg) * 16 + h * 8 + g;
            if (active && row < p.out_rows) {
                double* wrow = (splitk ? p.partial + (size_t)part * p.out_rows * p.ldw : p.W) + row * p.ldw + nb0 * 8 + 2 * t;
                const double sc = splitk ? 1.0 : p.alpha;
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    if (nb0 + nb < NB) {
                        double2 v;
                        v.x = sc * acc[h][nb][0];
                        v.y = sc * acc[h][nb][1];
                        *reinterpret_cast<double2*>(wrow + nb * 8) = v;
                    }
                }
            }
        }
    }
}

// W = alpha * (partial[0] + partial[1] + ...) over the first `cols` columns (ascending part order: deterministic)
__global__ void dense_splitk_fixup_kernel(const double* __restrict__ partial, int ksplit, int64_t out_rows, int64_t ldw,
                                          int cols, double alpha, double* __restrict__ W) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = idx / cols;
    const int c = (int)(idx - row * cols);
    if (row >= out_rows) return;
    double s = 0.0;
    for (int q = 0; q < ksplit; ++q) s += partial[((size_t)q * out_rows + row) * ldw + c];
    W[row * ldw + c] = alpha * s;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        GSI_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
        GSI_REQUIRE(ptr != nullptr && qres == cudaDriverEntryPointSuccess, GSI_ERR_CUDA,
                    "cuTensorMapEncodeTiled entry point not available");
        fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

static CUtensorMap make_map(const gsi_buf* A, int box_inner, int box_outer) {
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)A->rows, (cuuint64_t)A->cols};
    cuuint64_t strides[1] = {(cuuint64_t)A->ld * sizeof(double)};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)A->d, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    GSI_REQUIRE(r == CUDA_SUCCESS, GSI_ERR_CUDA, "cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
    return map;
}

template <int NB, int TRANS>
static void launch_dense(gsi_ctx* ctx, const gsi_buf* A, DenseParams p) {
    constexpr int ld = NB * 8 + 4;
    constexpr int a_inner = (TRANS ? DG_BK : DG_BM) + DG_PAD;
    constexpr int a_outer = TRANS ? DG_BM : DG_BK;
    constexpr size_t stage_bytes = (size_t)(DG_BK * ld + a_inner * a_outer) * sizeof(double);
    int stages = (int)((232448 - 256) / stage_bytes);      // all 227 KB a CTA may use: 3 stages up to 224 columns
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 2 * stages * sizeof(uint64_t);
    CUtensorMap map = make_map(A, a_inner, a_outer);
    auto kfn = dense_gemm_kernel<NB, TRANS>;
    GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, DG_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t ntiles = (p.out_rows + DG_BM - 1) / DG_BM;
    const int64_t nkt = (p.kdim + DG_BK - 1) / DG_BK;
    const int64_t ncta = (int64_t)ctx->num_sms * occ;
    int64_t grid = ncta;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    // split-K when the output has too few tiles to occupy the device and the reduction is long
    p.ksplit = 1; p.kt_per = nkt; p.partial = nullptr;
    size_t part_bytes = 0;
    if (ntiles * 2 <= ncta && nkt >= 32) {
        int64_t ks = ncta / ntiles;
        if (ks > nkt / 8) ks = nkt / 8;                     // at least 8 k-tiles per part
        if (ks > 1) {
            p.kt_per = (nkt + ks - 1) / ks;
            ks = (nkt + p.kt_per - 1) / p.kt_per;           // every part owns at least one k-tile
            p.ksplit = (int)ks;
            grid = ntiles * ks;
            part_bytes = (size_t)ks * p.out_rows * p.ldw * sizeof(double);
            p.partial = static_cast<double*>(pool_alloc(ctx, part_bytes));
        }
    }
    struct PartGuard { gsi_ctx* c; void* q; size_t b; ~PartGuard() { if (q) pool_free(c, q, b); } } pguard{ctx, p.partial, part_bytes};
    kfn<<<(unsigned)grid, DG_THREADS, smem, ctx->stream>>>(map, p);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
    if (p.ksplit > 1) {
        const int64_t total = p.out_rows * (8 * NB);
        dense_splitk_fixup_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p.partial, p.ksplit, p.out_rows, p.ldw,
                                                                                           8 * NB, p.alpha, p.W);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
}

void dense_apply(gsi_ctx* ctx, const gsi_buf* A, int trans, const gsi_buf* X, gsi_buf* W, double alpha) {
    GSI_REQUIRE(A->layout == GSI_LAYOUT_COLMAJOR, GSI_ERR_INVALID_ARGUMENT, "dense apply: A must be COLMAJOR");
    GSI_REQUIRE(X->layout == GSI_LAYOUT_TALL && W->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT,
                "dense apply: X and W must be TALL");
    GSI_REQUIRE(X->cols == W->cols, GSI_ERR_DIMENSION_MISMATCH, "dense apply: X/W column mismatch");
    if (X->cols > kMaxCols) {
        // wide iterate (K + p > 256): the product runs on 256-column chunks packed into compact buffers
        for (int64_t c0 = 0; c0 < X->cols; c0 += kMaxCols) {
            const int64_t w = (X->cols - c0 < kMaxCols) ? X->cols - c0 : kMaxCols;
            BufPtr xc = make_buf(ctx, GSI_LAYOUT_TALL, X->rows, w);
            BufPtr wc = make_buf(ctx, GSI_LAYOUT_TALL, W->rows, w);
            tall_cols_copy(ctx, X->d + c0, X->ld, xc->d, xc->ld, X->rows, w);
            dense_apply(ctx, A, trans, xc.get(), wc.get(), alpha);
            tall_cols_copy(ctx, wc->d, wc->ld, W->d + c0, W->ld, W->rows, w);
        }
        return;
    }
    const int nb = nb_for_cols(X->cols);
    GSI_REQUIRE(X->ld == 8 * nb + 4 && W->ld == X->ld, GSI_ERR_INVALID_ARGUMENT, "dense apply: bad pitch");
    DenseParams p;
    p.X = X->d; p.W = W->d; p.ld = X->ld; p.ldw = W->ld; p.alpha = alpha; p.stages = 0;
    if (!trans) {
        GSI_REQUIRE(X->rows == A->cols, GSI_ERR_DIMENSION_MISMATCH, "dense apply: A*X inner dimension");
        GSI_REQUIRE(W->rows == A->rows, GSI_ERR_DIMENSION_MISMATCH, "dense apply: A*X output rows");
        p.out_rows = A->rows; p.kdim = A->cols;
    } else {
        GSI_REQUIRE(X->rows == A->rows, GSI_ERR_DIMENSION_MISMATCH, "dense apply: A'*X inner dimension");
        GSI_REQUIRE(W->rows == A->cols, GSI_ERR_DIMENSION_MISMATCH, "dense apply: A'*X output rows");
        p.out_rows = A->cols; p.kdim = A->rows;
    }
    switch (nb) {
#define GSI_CASE(N) case N: if (trans) launch_dense<N, 1>(ctx, A, p); else launch_dense<N, 0>(ctx, A, p); break;
        GSI_NB_LIST(GSI_CASE)
#undef GSI_CASE
        default: throw Error(GSI_ERR_UNSUPPORTED, "dense apply: unsupported column-block count");
    }
}

}  // namespace gsi
