// Matrix-free covariance-kernel operator  W[I, :] = C[I, :] * X   (SURVEY.md §8 a8).
//
// C[i,j] = sigma2 * k(r2(u_i, u_j)) + nugget * (i == j) is never materialised: every
// lane generates its one entry of an 8x4 DMMA A-fragment directly in registers
// (row = lane/4, column = lane%4), multiplies it against the X tile staged in shared
// memory by 1-D bulk async copies (UBLKCP, mbarrier completion), and accumulates an
// 8 x (8*NB) output strip per warp in registers with FP64 tensor-core MMAs
// (mma.sync.m8n8k4.f64 -> DMMA.8x8x4).
//
// CTA = 8 warps (64 output rows x all 8*NB columns); thread 0 also issues the bulk
// copies `lookahead` k-tiles ahead of the math (a 9th producer warp would make ptxas
// budget registers for 12 warps -> 168/thread, spilling the 27-block accumulator).
// Persistent grid: one CTA per SM (x occupancy), static round-robin over row tiles.
// X lives in the TALL layout (row pitch ld = 8*NB + 4 doubles), so a BK-row tile is one
// contiguous chunk and the B-fragment LDS.64 pattern (4 rows x 4 col-octets per
// half-warp) is bank-conflict free.
#include "common.cuh"
#include "ptx.cuh"
#include "nb_list.h"

namespace gsi {

constexpr int KC_BM = 64;          // rows per CTA tile (8 warps x 8 rows)
constexpr int KC_BK = 32;          // j-points (GEMM K) per pipeline stage
constexpr int KC_CONSUMERS = 8;
constexpr int KC_THREADS = KC_CONSUMERS * 32;   // thread 0 doubles as the bulk-copy producer

struct KcovParams {
    const double* X;       // TALL, all n rows (zero padded), pitch ld
    double* W;             // TALL, local rows, pitch ldw
    const double* u;       // scaled coordinates [3][n_pad]
    int64_t n;             // columns of C (= rows of X)
    int64_t n_pad;
    int64_t row0;          // first global row of this rank's block
    int64_t mloc;          // local rows
    int64_t ld, ldw;
    double sigma2, nugget, beta;
    int stages;
};

template <int KIND>
__device__ __forceinline__ double kern_eval(double r2, double beta) {
    if (KIND == GSI_KERNEL_EXPONENTIAL) return exp(-sqrt(r2));
    if (KIND == GSI_KERNEL_GAUSSIAN) return exp(-0.5 * r2);
    return exp(-beta * log1p(r2));
}

template <int NB, int KIND, int DIM>
__global__ void __launch_bounds__(KC_THREADS, 1) kcov_gemm_kernel(const __grid_constant__ KcovParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int ld = NB * 8 + 4;
    const int stage_doubles = KC_BK * ld + 3 * KC_BK;          // X tile + coordinate tile
    double* smem = reinterpret_cast<double*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_doubles);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nstages = p.stages;

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], KC_CONSUMERS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t ntiles = (p.mloc + KC_BM - 1) / KC_BM;
    const int64_t nkt = (p.n + KC_BK - 1) / KC_BK;
    const uint32_t stage_bytes = (uint32_t)((KC_BK * ld + DIM * KC_BK) * sizeof(double));

    // ---------------- producer (thread 0): streams X / coordinate tiles ----------------------
    const int64_t my_tiles = (ntiles > blockIdx.x) ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total_it = my_tiles * nkt;
    const int lookahead = nstages > 2 ? nstages - 2 : 1;
    auto produce = [&](int64_t nxt) {
        const int s = (int)(nxt % nstages);
        const uint32_t ph = (uint32_t)((nxt / nstages) & 1);
        const int64_t kt = nxt % nkt;
        mbar_wait(&empty[s], ph ^ 1u);
        double* xs = smem + (size_t)s * stage_doubles;
        double* us = xs + KC_BK * ld;
        mbar_expect_tx(&full[s], stage_bytes);
        bulk_g2s(xs, p.X + kt * KC_BK * p.ld, KC_BK * ld * 8, &full[s]);
#pragma unroll
        for (int k = 0; k < DIM; ++k)
            bulk_g2s(us + k * KC_BK, p.u + k * p.n_pad + kt * KC_BK, KC_BK * 8, &full[s]);
    };
    if (tid == 0) {
        for (int64_t i = 0; i < lookahead && i < total_it; ++i) produce(i);
    }

    // ---------------- consumer warps -------------------------------------------------------
    const int g = lane >> 2;      // fragment row (A, C) / column (B)
    const int t = lane & 3;       // fragment k index (A, B) / column pair (C)
    int64_t it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t lrow = tile * KC_BM + warp * 8 + g;              // local output row
        int64_t grow = p.row0 + lrow;                                  // global point index
        if (grow > p.n - 1) grow = p.n - 1;                            // tail rows: clamp (never stored)
        double ui[DIM];
#pragma unroll
        for (int k = 0; k < DIM; ++k) ui[k] = p.u[k * p.n_pad + grow];

        double acc[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) { acc[nb][0] = 0.0; acc[nb][1] = 0.0; }

        for (int64_t kt = 0; kt < nkt; ++kt, ++it) {
            if (tid == 0 && it + lookahead < total_it) produce(it + lookahead);
            const int s = (int)(it % nstages);
            const uint32_t ph = (uint32_t)((it / nstages) & 1);
            mbar_wait(&full[s], ph);
            __syncwarp();
            const double* xs = smem + (size_t)s * stage_doubles;
            const double* us = xs + KC_BK * ld;
#pragma unroll 2
            for (int ks = 0; ks < KC_BK / 4; ++ks) {
                const int j = ks * 4 + t;
                double r2 = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; ++k) {
                    const double dk = ui[k] - us[k * KC_BK + j];
                    r2 += dk * dk;
                }
                const double a = kern_eval<KIND>(r2, p.beta);
                const double* xrow = xs + j * ld + g;
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) dmma884(acc[nb][0], acc[nb][1], a, xrow[nb * 8]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }

        // epilogue: W = sigma2 * acc + nugget * X[global row]
        if (lrow < p.mloc) {
            double* wrow = p.W + lrow * p.ldw + 2 * t;
            const double* xg = p.X + (p.row0 + lrow) * p.ld + 2 * t;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) {
                double2 v;
                v.x = p.sigma2 * acc[nb][0];
                v.y = p.sigma2 * acc[nb][1];
                if (p.nugget != 0.0) {
                    const double2 xv = *reinterpret_cast<const double2*>(xg + nb * 8);
                    v.x += p.nugget * xv.x;
                    v.y += p.nugget * xv.y;
                }
                *reinterpret_cast<double2*>(wrow + nb * 8) = v;
            }
        }
    }
}

template <int NB, int KIND, int DIM>
static void launch_kcov(gsi_ctx* ctx, const KcovParams& p0) {
    KcovParams p = p0;
    const int ld = NB * 8 + 4;
    const size_t stage_bytes = (size_t)(KC_BK * ld + 3 * KC_BK) * sizeof(double);
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + 2 * stages * sizeof(uint64_t);
    auto kfn = kcov_gemm_kernel<NB, KIND, DIM>;
    GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, KC_THREADS, smem));
    if (occ < 1) occ = 1;
    const int64_t ntiles = (p.mloc + KC_BM - 1) / KC_BM;
    int64_t grid = (int64_t)ctx->num_sms * occ;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    kfn<<<(unsigned)grid, KC_THREADS, smem, ctx->stream>>>(p);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

template <int NB, int KIND>
static void dispatch_dim(gsi_ctx* ctx, const KcovParams& p, int dim) {
    if (dim <= 2) launch_kcov<NB, KIND, 2>(ctx, p);
    else launch_kcov<NB, KIND, 3>(ctx, p);
}

template <int NB>
static void dispatch_kind(gsi_ctx* ctx, const KcovParams& p, int kind, int dim) {
    switch (kind) {
        case GSI_KERNEL_EXPONENTIAL: dispatch_dim<NB, GSI_KERNEL_EXPONENTIAL>(ctx, p, dim); break;
        case GSI_KERNEL_GAUSSIAN: dispatch_dim<NB, GSI_KERNEL_GAUSSIAN>(ctx, p, dim); break;
        case GSI_KERNEL_POWERLAW: dispatch_dim<NB, GSI_KERNEL_POWERLAW>(ctx, p, dim); break;
        default: throw Error(GSI_ERR_INVALID_ARGUMENT, "unknown covariance kernel kind");
    }
}

void kcov_apply(gsi_op* op, const gsi_buf* X, gsi_buf* W) {
    gsi_ctx* ctx = op->ctx;
    GSI_REQUIRE(X->layout == GSI_LAYOUT_TALL && W->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT,
                "kernelcov apply needs TALL buffers");
    GSI_REQUIRE(X->rows == op->n, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: X must have n rows");
    GSI_REQUIRE(W->rows == op->mloc, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: W must have mloc rows");
    GSI_REQUIRE(X->cols == W->cols, GSI_ERR_DIMENSION_MISMATCH, "kernelcov apply: X/W column mismatch");
    GSI_REQUIRE(X->cols <= kMaxCols, GSI_ERR_UNSUPPORTED, "kernelcov apply: more than 256 columns");
    const int nb = nb_for_cols(X->cols);
    GSI_REQUIRE(X->ld == 8 * nb + 4 && W->ld == X->ld, GSI_ERR_INVALID_ARGUMENT, "kernelcov apply: bad pitch");
    KcovParams p;
    p.X = X->d; p.W = W->d; p.u = op->ucoords;
    p.n = op->n; p.n_pad = op->n_pad; p.row0 = op->row0; p.mloc = op->mloc;
    p.ld = X->ld; p.ldw = W->ld;
    p.sigma2 = op->sigma2; p.nugget = op->nugget; p.beta = op->beta;
    p.stages = 0;
    switch (nb) {
#define GSI_CASE(N) case N: dispatch_kind<N>(ctx, p, op->kind, op->dim); break;
        GSI_NB_LIST(GSI_CASE)
#undef GSI_CASE
        default: throw Error(GSI_ERR_UNSUPPORTED, "kernelcov apply: unsupported column-block count");
    }
}

}  // namespace gsi
