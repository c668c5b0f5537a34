// Matrix-free covariance-kernel operator  W[I, :] = C[I, :] * X   (SURVEY.md §8 a8).
//
// C[i,j] = sigma2 * k(r2(u_i, u_j)) + nugget * (i == j) is never materialised.  Per
// 32-point k-tile, the four warps that share a 16-row group generate that group's
// 16 x 32 block of kernel values ONCE (4 entries per thread, FP64 DFMA/exp) into a
// double-buffered shared-memory tile laid out for conflict-free DMMA A-fragment loads,
// synchronise on a 128-thread named barrier, and then each warp multiplies it against its
// quarter of the X tile with FP64 tensor-core MMAs (mma.sync.m8n8k4.f64 -> DMMA.8x8x4),
// accumulating a 16 x (8*NB/4) strip in registers.  X tiles (TALL layout, a k-tile is one
// contiguous chunk) and the scaled coordinates arrive by 1-D bulk async copies (UBLKCP)
// completing on mbarriers; thread 0 issues them `lookahead` tiles ahead.
//
// Why this shape (profiles/r01): a single warp can issue a DMMA only every ~32 cycles
// while the pipe retires one per 16, and DFMA shares that pipe -- so the SM needs >= 3-4
// warps per scheduler (16 warps/SM, <= 128 registers each) and every kernel value must be
// generated exactly once per CTA.  The 8-warp / A-in-registers first version stalled 43 %
// of samples in `wait` with the DMMA pipe 71 % busy.
//
// CTA = 16 warps = 4 row groups (16 rows) x 4 column groups: 64 rows x 8*NB columns.
// Persistent grid: one CTA per SM (x occupancy), static round-robin over row tiles.
#include "common.cuh"
#include "ptx.cuh"
#include "nb_list.h"

namespace gsi {

constexpr int KC_BM = 64;          // rows per CTA tile
constexpr int KC_BK = 32;          // j-points (GEMM K) per pipeline stage
constexpr int KC_AP = KC_BK + 4;   // pitch of the generated A tile (4 mod 16 doubles)
constexpr int KC_WARPS = 16;
constexpr int KC_THREADS = KC_WARPS * 32;
constexpr int KC_CG = 4;           // column groups (warps sharing one row group)

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

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

template <int NB, int KIND, int DIM>
__global__ void __launch_bounds__(KC_THREADS, 1) kcov_gemm_kernel(const __grid_constant__ KcovParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ld = NB * 8 + 4;
    constexpr int NBW = (NB + KC_CG - 1) / KC_CG;                // n-blocks per warp
    constexpr int stage_doubles = KC_BK * ld + 3 * KC_BK;        // X tile + coordinate tile
    constexpr int a_doubles = KC_BM * KC_AP;
    double* smem = reinterpret_cast<double*>(smem_raw);
    double* a_tiles = smem + (size_t)p.stages * stage_doubles;   // [2][KC_BM][KC_AP]
    uint64_t* full = reinterpret_cast<uint64_t*>(a_tiles + 2 * a_doubles);
    uint64_t* empty = full + p.stages;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nstages = p.stages;
    const int rg = warp >> 2;                 // row group: rows rg*16 .. rg*16+15 of the tile
    const int cg = (warp + rg) & 3;           // column group, rotated per row group so that each SM
                                              // sub-partition (warp % 4) hosts all four column groups
                                              // (the last group may own fewer n-blocks)
    const int nb0 = cg * NBW;                 // first n-block of this warp

    if (tid == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], KC_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t ntiles = (p.mloc + KC_BM - 1) / KC_BM;
    const int64_t nkt = (p.n + KC_BK - 1) / KC_BK;
    constexpr uint32_t stage_bytes = (uint32_t)((KC_BK * ld + DIM * KC_BK) * sizeof(double));

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
        const int64_t ktn = (kt + 1 == nkt) ? 0 : kt + 1;      // coordinates of the NEXT k-tile ride along
        for (int k = 0; k < DIM; ++k)
            bulk_g2s(us + k * KC_BK, p.u + k * p.n_pad + ktn * KC_BK, KC_BK * 8, &full[s]);
    };
    if (tid == 0) {
        for (int64_t i = 0; i < lookahead && i < total_it; ++i) produce(i);
    }

    const int g = lane >> 2;      // fragment row (A, C) / column (B)
    const int t = lane & 3;       // fragment k index (A, B) / column pair (C)
    // generation role inside the row group: 128 threads x 4 entries = 16 rows x 32 points
    const int gtid = tid & 127;                  // thread index within the row group
    const int grow_in_tile = rg * 16 + (gtid >> 3);
    const int gj0 = (gtid & 7) * 4;
    auto row_coords = [&](int64_t tile, double (&ui)[DIM]) {
        int64_t gpt = p.row0 + tile * KC_BM + grow_in_tile;           // global point of my generated row
        if (gpt > p.n - 1) gpt = p.n - 1;                             // tail rows: clamp (never stored)
#pragma unroll
        for (int k = 0; k < DIM; ++k) ui[k] = p.u[k * p.n_pad + gpt];
    };
    auto gen4 = [&](const double (&ui)[DIM], const double* u0, int64_t ustride, double (&v)[4]) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            double r2 = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) {
                const double dk = ui[k] - u0[k * ustride + gj0 + e];
                r2 += dk * dk;
            }
            v[e] = kern_eval<KIND>(r2, p.beta);
        }
    };
    auto store4 = [&](double* as, const double (&v)[4]) {
        double2* dst = reinterpret_cast<double2*>(as + grow_in_tile * KC_AP + gj0);
        dst[0] = make_double2(v[0], v[1]);
        dst[1] = make_double2(v[2], v[3]);
    };

    int64_t it = 0;
    double ui[DIM];
    if (blockIdx.x < ntiles) {
        // prologue: kernel values of (first tile, k-tile 0) straight from global coordinates
        row_coords(blockIdx.x, ui);
        double v[4];
        gen4(ui, p.u, p.n_pad, v);
        store4(a_tiles, v);
        named_bar_sync(1 + rg, 128);
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        double acc[2][NBW][2];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0.0; acc[h][nb][1] = 0.0; }

        for (int64_t kt = 0; kt < nkt; ++kt, ++it) {
            if (tid == 0 && it + lookahead < total_it) produce(it + lookahead);
            const int s = (int)(it % nstages);
            const uint32_t ph = (uint32_t)((it / nstages) & 1);
            mbar_wait(&full[s], ph);
            __syncwarp();
            const double* xs = smem + (size_t)s * stage_doubles;
            const double* us = xs + KC_BK * ld;                       // coordinates of k-tile kt+1 (wraps to 0)
            const double* as = a_tiles + (size_t)(it & 1) * a_doubles;
            // ---- kernel values of the NEXT k-tile (independent of the MMAs below: the two
            //      instruction streams overlap on the shared FP64 pipe)
            if (kt + 1 == nkt) row_coords(tile + gridDim.x < ntiles ? tile + gridDim.x : tile, ui);
            double vnext[4];
            gen4(ui, us, KC_BK, vnext);
            // ---- tensor-core phase on the current k-tile
            const double* arow0 = as + (rg * 16 + g) * KC_AP + t;
            const double* arow1 = arow0 + 8 * KC_AP;
#pragma unroll 2
            for (int ks = 0; ks < KC_BK / 4; ++ks) {
                const double a0 = arow0[ks * 4];
                const double a1 = arow1[ks * 4];
                const double* xrow = xs + (ks * 4 + t) * ld + nb0 * 8 + g;
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    if (nb0 + nb < NB) {
                        const double b = xrow[nb * 8];
                        dmma884(acc[0][nb][0], acc[0][nb][1], a0, b);
                        dmma884(acc[1][nb][0], acc[1][nb][1], a1, b);
                    }
                }
            }
            store4(a_tiles + (size_t)((it + 1) & 1) * a_doubles, vnext);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            named_bar_sync(1 + rg, 128);          // next block complete; current block no longer read
        }

        // epilogue: W = sigma2 * acc + nugget * X[global row]
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t lrow = tile * KC_BM + rg * 16 + h * 8 + g;
            if (lrow < p.mloc) {
                double* wrow = p.W + lrow * p.ldw + nb0 * 8 + 2 * t;
                const double* xg = p.X + (p.row0 + lrow) * p.ld + nb0 * 8 + 2 * t;
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    if (nb0 + nb < NB) {
                        double2 v;
                        v.x = p.sigma2 * acc[h][nb][0];
                        v.y = p.sigma2 * acc[h][nb][1];
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
    }
}

template <int NB, int KIND, int DIM>
static void launch_kcov(gsi_ctx* ctx, const KcovParams& p0) {
    KcovParams p = p0;
    const int ld = NB * 8 + 4;
    const size_t stage_bytes = (size_t)(KC_BK * ld + 3 * KC_BK) * sizeof(double);
    const size_t a_bytes = (size_t)2 * KC_BM * KC_AP * sizeof(double);
    int stages = (int)((224 * 1024 - a_bytes - 256) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = stages * stage_bytes + a_bytes + 2 * stages * sizeof(uint64_t);
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
