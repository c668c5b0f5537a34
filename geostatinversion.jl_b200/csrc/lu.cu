// Reference-faithful normaliser:  `F = lu(Y); Q = F.L`
// (reference src/RandMatFact.jl:60-61,68-69,72-73 -> LAPACK dgetrf).
//
// Gaussian elimination with partial pivoting on a tall n x l TALL buffer, in place,
// with LAPACK's pivot rule (idamax: FIRST row of maximal |value|), physical row
// interchanges and multiplication by the reciprocal pivot (dgetf2/dgetrf2).  The result
// is the unit-lower-trapezoidal factor in LAPACK's *permuted* row order -- the reference
// never un-permutes it.
//
// Blocked right-looking algorithm (panel width LU_PB): inside a panel only the panel's own
// columns are eliminated column by column (the n x 16 panel stays L2-resident); after the
// panel, U12 = L11^{-1} A12 is formed by one small kernel and the trailing matrix receives
// ONE rank-LU_PB update on the FP64 tensor cores (tall_window_update -> dense DMMA GEMM),
// i.e. it is read and written l/LU_PB times instead of l times.  Row interchanges always
// move whole rows, as LAPACK's dlaswp does.
//
// One column = two launches on the context stream (no host sync):
//   lu_pack      (1 CTA)  reduce the per-CTA pivot candidates of column k, export the
//                         candidate row and (if owned) row k            -> send buffer
//   [allgather over ranks when the rows are sharded]
//   lu_eliminate (grid)   pick the global pivot, move rows k <-> p from the exchanged
//                         copies (never from memory being rewritten), scale column k,
//                         rank-1 update of the trailing columns, and -- fused -- the
//                         arg-max search of column k+1.
//
// EXPERIMENTAL (option "lu.fused", default off, single-GPU / replicated iterates only): the
// column steps of a panel in ONE cooperative launch, lu_panel_fused_kernel -- the same
// arithmetic per element, one grid-wide barrier per column instead of two kernel boundaries.
// Every row is written only by the CTA that owns it; the pivot candidates travel with a copy of
// their row, so nobody reads a row another CTA may be rewriting.  Not yet run on hardware.
#include "common.cuh"
#include "algos.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gsi {

constexpr int LU_PB = 16;            // panel width
constexpr int LU_THREADS = 256;
constexpr int LU_WARPS = LU_THREADS / 32;

struct Cand { double val; double idx; };   // idx = global row index (exact in a double)

__device__ __forceinline__ bool cand_better(double v, double i, double bv, double bi) {
    return (v > bv) || (v == bv && i < bi);
}

// column-0 search (later columns are searched inside lu_eliminate)
__global__ void lu_search_kernel(const double* __restrict__ Y, int64_t ld, int64_t nloc, int64_t row0, int col,
                                 Cand* __restrict__ cand) {
    double bv = -1.0, bi = 0.0;
    const int64_t gstart = col;   // rows with global index >= col
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gi = row0 + i;
        if (gi < gstart) continue;
        const double v = fabs(Y[i * ld + col]);
        if (cand_better(v, (double)gi, bv, bi)) { bv = v; bi = (double)gi; }
    }
    __shared__ double sv[LU_THREADS], si[LU_THREADS];
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = LU_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            if (cand_better(sv[threadIdx.x + s], si[threadIdx.x + s], sv[threadIdx.x], si[threadIdx.x])) {
                sv[threadIdx.x] = sv[threadIdx.x + s]; si[threadIdx.x] = si[threadIdx.x + s];
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { cand[blockIdx.x].val = sv[0]; cand[blockIdx.x].idx = si[0]; }
}

// send layout: [0] = |candidate|, (-1 if this rank has no row >= k), [1] = global row,
//              [2 .. 2+l) candidate row, [2+l .. 2+2l) row k (only meaningful on its owner)
__global__ void lu_pack_kernel(const double* __restrict__ Y, int64_t ld, int64_t nloc, int64_t row0, int l, int k,
                               const Cand* __restrict__ cand, int ncand, double* __restrict__ send) {
    __shared__ double sv[LU_THREADS], si[LU_THREADS];
    double bv = -1.0, bi = 0.0;
    for (int c = threadIdx.x; c < ncand; c += LU_THREADS) {
        const double v = cand[c].val, i = cand[c].idx;
        if (v >= 0.0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; }
    }
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = LU_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            if (cand_better(sv[threadIdx.x + s], si[threadIdx.x + s], sv[threadIdx.x], si[threadIdx.x])) {
                sv[threadIdx.x] = sv[threadIdx.x + s]; si[threadIdx.x] = si[threadIdx.x + s];
            }
        }
        __syncthreads();
    }
    bv = sv[0]; bi = si[0];
    if (threadIdx.x == 0) { send[0] = bv; send[1] = bi; }
    if (bv >= 0.0) {
        const int64_t li = (int64_t)bi - row0;
        for (int j = threadIdx.x; j < l; j += LU_THREADS) send[2 + j] = Y[li * ld + j];
    }
    const int64_t lk = (int64_t)k - row0;
    if (lk >= 0 && lk < nloc)
        for (int j = threadIdx.x; j < l; j += LU_THREADS) send[2 + l + j] = Y[lk * ld + j];
}

// Elimination of column k restricted to the panel columns (k, jend).  8 rows per warp,
// 4 lanes per row (a panel row segment is <= 15 contiguous doubles).
__global__ void __launch_bounds__(LU_THREADS)
lu_eliminate_kernel(double* __restrict__ Y, int64_t ld, int64_t nloc, int64_t row0, int l, int k, int jend,
                    const double* __restrict__ recv, int world, int owner_k, Cand* __restrict__ cand,
                    int* __restrict__ flags) {
    extern __shared__ double sm[];
    double* prow = sm;            // pivot row (all l columns)
    double* krow = sm + l;        // previous content of row k
    __shared__ double s_best[LU_WARPS], s_bidx[LU_WARPS];
    __shared__ double s_p;
    const int stride = 2 + 2 * l;
    if (threadIdx.x == 0) {
        double bv = -1.0, bi = 0.0;
        int win = 0;
        for (int r = 0; r < world; ++r) {
            const double v = recv[(size_t)r * stride], i = recv[(size_t)r * stride + 1];
            if (v >= 0.0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; win = r; }
        }
        s_p = bi;
        s_best[0] = (double)win;
    }
    __syncthreads();
    const int win = (int)s_best[0];
    const int64_t p = (int64_t)s_p;
    __syncthreads();
    for (int j = threadIdx.x; j < l; j += LU_THREADS) {
        prow[j] = recv[(size_t)win * stride + 2 + j];
        krow[j] = recv[(size_t)owner_k * stride + 2 + l + j];
    }
    __syncthreads();
    const double pivot = prow[k];
    double rpiv = 0.0;
    if (pivot == 0.0) {
        if (blockIdx.x == 0 && threadIdx.x == 0) atomicCAS(&flags[0], 0, k + 1);   // first zero pivot (1-based)
    } else {
        rpiv = 1.0 / pivot;
    }
    // LAPACK dgetf2: reciprocal scaling when |pivot| >= sfmin, true division otherwise
    const bool use_recip = fabs(pivot) >= 2.2250738585072014e-308;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && p != k) {
        // row k receives the pivot row (all columns); the part of old row k that this step does
        // not rewrite (L part and columns beyond the panel) moves to position p
        const int64_t lk = (int64_t)k - row0, lp = p - row0;
        if (lk >= 0 && lk < nloc)
            for (int j = threadIdx.x; j < l; j += LU_THREADS) Y[lk * ld + j] = prow[j];
        if (lp >= 0 && lp < nloc)
            for (int j = threadIdx.x; j < l; j += LU_THREADS)
                if (j < k || j >= jend) Y[lp * ld + j] = krow[j];
    }

    const int sub = lane & 3, rsub = lane >> 2;
    double bv = -1.0, bi = 0.0;
    int64_t lstart = (int64_t)k + 1 - row0;
    if (lstart < 0) lstart = 0;
    const int64_t wglobal = (int64_t)blockIdx.x * LU_WARPS + warp;
    const int64_t wtotal = (int64_t)gridDim.x * LU_WARPS;
    for (int64_t i = lstart + wglobal * 8 + rsub; i < nloc; i += wtotal * 8) {
        const int64_t gi = row0 + i;
        double* yrow = Y + i * ld;
        const double* src = (gi == p) ? krow : yrow;
        double m;
        if (pivot == 0.0) m = src[k];
        else m = use_recip ? src[k] * rpiv : src[k] / pivot;
        if (sub == 0) yrow[k] = m;
        for (int j = k + 1 + sub; j < jend; j += 4) {
            const double v = src[j] - m * prow[j];
            yrow[j] = v;
            if (j == k + 1) {   // sub == 0: candidate for the next column of this panel
                const double a = fabs(v);
                if (cand_better(a, (double)gi, bv, bi)) { bv = a; bi = (double)gi; }
            }
        }
    }
    // candidates live in lanes with sub == 0: reduce over the 8 row slots of the warp
    for (int o = 4; o < 32; o <<= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o), oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { s_best[warp] = bv; s_bidx[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < LU_WARPS; ++w)
            if (cand_better(s_best[w], s_bidx[w], bv, bi)) { bv = s_best[w]; bi = s_bidx[w]; }
        cand[blockIdx.x].val = bv; cand[blockIdx.x].idx = bi;
    }
}


// ---------------------------------------------------------------------------- fused panel
// One cooperative launch per panel [ps, pe).  CTA b owns rows [b*R, (b+1)*R) (R a multiple of
// 8).  Buffers (double-buffered by column parity so that a CTA that is already publishing for
// column k+1 cannot overwrite what a slower CTA still reads for column k; a buffer of a given
// parity is rewritten only after two grid barriers):
//   cand [2][G]      best (|value|, row) of each CTA for the current column
//   rows [2][G][l]   copy of that candidate row (all l columns)
//   krows[2][l]      copy of row k, published by its owner
struct LuFusedParams {
    double* Y; int64_t ld; int64_t n; int l; int ps, pe;
    Cand* cand; double* rows; double* krows; int* flags;
    int64_t R;
};

// (|v|, row) arg-max over the block with LAPACK's tie rule; every thread gets the result.
__device__ void lu_block_best(double& bv, double& bi, double* sv, double* si) {
    sv[threadIdx.x] = bv; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = LU_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            if (cand_better(sv[threadIdx.x + s], si[threadIdx.x + s], sv[threadIdx.x], si[threadIdx.x])) {
                sv[threadIdx.x] = sv[threadIdx.x + s]; si[threadIdx.x] = si[threadIdx.x + s];
            }
        }
        __syncthreads();
    }
    bv = sv[0]; bi = si[0];
    __syncthreads();
}

__global__ void __launch_bounds__(LU_THREADS) lu_panel_fused_kernel(LuFusedParams p) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double sm[];
    double* prow = sm;              // pivot row (all l columns)
    double* krow = sm + p.l;        // previous content of row k
    __shared__ double sv[LU_THREADS], si[LU_THREADS];
    const int l = p.l, G = (int)gridDim.x, b = (int)blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane & 3, rsub = lane >> 2;
    const int64_t r0 = (int64_t)b * p.R;
    const int64_t r1 = (r0 + p.R < p.n) ? r0 + p.R : p.n;      // my rows [r0, r1) (may be empty)
    double* Y = p.Y;
    const int64_t ld = p.ld;

    // publish my candidate (and its row) for column `col`, and row `col` if I own it
    auto publish = [&](int col, int par, double bv, double bi) {
        lu_block_best(bv, bi, sv, si);
        if (threadIdx.x == 0) { p.cand[(size_t)par * G + b].val = bv; p.cand[(size_t)par * G + b].idx = bi; }
        if (bv >= 0.0) {
            const double* src = Y + (int64_t)bi * ld;
            double* dst = p.rows + ((size_t)par * G + b) * l;
            for (int j = threadIdx.x; j < l; j += LU_THREADS) dst[j] = src[j];
        }
        if (col >= r0 && col < r1) {
            const double* src = Y + (int64_t)col * ld;
            double* dst = p.krows + (size_t)par * l;
            for (int j = threadIdx.x; j < l; j += LU_THREADS) dst[j] = src[j];
        }
    };

    // ---- candidates of the first column of the panel
    {
        double bv = -1.0, bi = 0.0;
        for (int64_t i = r0 + threadIdx.x; i < r1; i += LU_THREADS) {
            if (i < p.ps) continue;
            const double v = fabs(Y[i * ld + p.ps]);
            if (cand_better(v, (double)i, bv, bi)) { bv = v; bi = (double)i; }
        }
        publish(p.ps, 0, bv, bi);
    }
    grid.sync();

    for (int k = p.ps; k < p.pe; ++k) {
        const int par = (k - p.ps) & 1;
        // ---- global pivot: every CTA reduces the G candidates the same way
        double bv = -1.0, bi = 0.0;
        for (int c = threadIdx.x; c < G; c += LU_THREADS) {
            const double v = p.cand[(size_t)par * G + c].val, i = p.cand[(size_t)par * G + c].idx;
            if (v >= 0.0 && cand_better(v, i, bv, bi)) { bv = v; bi = i; }
        }
        lu_block_best(bv, bi, sv, si);
        const int64_t piv = (int64_t)bi;
        const int win = (int)(piv / p.R);                       // the CTA that owns (and published) row piv
        for (int j = threadIdx.x; j < l; j += LU_THREADS) {
            prow[j] = p.rows[((size_t)par * G + win) * l + j];
            krow[j] = p.krows[(size_t)par * l + j];
        }
        __syncthreads();
        const double pivot = prow[k];
        double rpiv = 0.0;
        if (pivot == 0.0) {
            if (b == 0 && threadIdx.x == 0) atomicCAS(&p.flags[0], 0, k + 1);   // first zero pivot (1-based)
        } else {
            rpiv = 1.0 / pivot;
        }
        const bool use_recip = fabs(pivot) >= 2.2250738585072014e-308;       // dgetf2: sfmin
        // ---- row interchange, each row by its owner: row k receives the pivot row (all columns);
        //      the part of old row k that this step does not rewrite moves to position piv
        if (piv != k) {
            if (k >= r0 && k < r1)
                for (int j = threadIdx.x; j < l; j += LU_THREADS) Y[(int64_t)k * ld + j] = prow[j];
            if (piv >= r0 && piv < r1)
                for (int j = threadIdx.x; j < l; j += LU_THREADS)
                    if (j < k || j >= p.pe) Y[piv * ld + j] = krow[j];
        }
        // ---- elimination of my rows below k, candidates for column k+1
        bv = -1.0; bi = 0.0;
        int64_t lstart = (int64_t)k + 1;
        if (lstart < r0) lstart = r0;
        for (int64_t i = lstart + warp * 8 + rsub; i < r1; i += LU_WARPS * 8) {
            double* yrow = Y + i * ld;
            const double* src = (i == piv) ? krow : yrow;
            double m;
            if (pivot == 0.0) m = src[k];
            else m = use_recip ? src[k] * rpiv : src[k] / pivot;
            if (sub == 0) yrow[k] = m;
            for (int j = k + 1 + sub; j < p.pe; j += 4) {
                const double v = src[j] - m * prow[j];
                yrow[j] = v;
                if (j == k + 1) {
                    const double a = fabs(v);
                    if (cand_better(a, (double)i, bv, bi)) { bv = a; bi = (double)i; }
                }
            }
        }
        if (k + 1 < p.pe) {
            __syncthreads();                     // my rows are complete before their copies are taken
            publish(k + 1, par ^ 1, bv, bi);
            grid.sync();
        }
    }
}

// U12 = L11^{-1} A12 for the panel rows [ps, pe): one thread per trailing column.  The rows are
// updated in place and copied to the small TALL buffer U (pb x (l - pe)) for the GEMM update.
__global__ void lu_u12_kernel(double* __restrict__ Y, int64_t ld, int l, int ps, int pe, int64_t lrow_ps,
                              double* __restrict__ U, int64_t ldu) {
    const int j = pe + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= l) return;
    double u[LU_PB];
    const int pb = pe - ps;
#pragma unroll
    for (int r = 0; r < LU_PB; ++r) {
        if (r < pb) {
            const double* yr = Y + (lrow_ps + r) * ld;
            double v = yr[j];
#pragma unroll
            for (int c = 0; c < LU_PB; ++c)
                if (c < r) v -= yr[ps + c] * u[c];
            u[r] = v;
            Y[(lrow_ps + r) * ld + j] = v;
            U[(int64_t)r * ldu + (j - pe)] = v;
        }
    }
}

// rows with global index < l: zero the U part, unit diagonal
__global__ void lu_finalize_kernel(double* __restrict__ Y, int64_t ld, int64_t nloc, int64_t row0, int l) {
    const int64_t gi = blockIdx.x;
    const int64_t i = gi - row0;
    if (i < 0 || i >= nloc) return;
    for (int j = (int)gi + threadIdx.x; j < l; j += blockDim.x) Y[i * ld + j] = (j == gi) ? 1.0 : 0.0;
}

void lu_L_inplace(gsi_ctx* ctx, gsi_buf* Y, int64_t row0, int64_t n_global, const int64_t* part_row0) {
    GSI_REQUIRE(Y->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "lu: TALL buffer required");
    const int l = (int)Y->cols;
    const int64_t nloc = Y->rows;
    GSI_REQUIRE(n_global >= l, GSI_ERR_UNSUPPORTED, "lu: fewer rows than columns is not supported");
    const int world = ctx->world;
    int grid = ctx->num_sms * 4;
    const int64_t need = (nloc + LU_WARPS * 8 - 1) / (LU_WARPS * 8);
    if (grid > need) grid = (int)(need > 0 ? need : 1);
    const size_t stride = 2 + 2 * (size_t)l;
    // scratch layout: cand[grid] | send[stride] | recv[world*stride]
    const size_t need_doubles = 2 * (size_t)grid + stride * (1 + (size_t)world) + 16;
    GSI_REQUIRE(need_doubles <= ctx->scratch_doubles, GSI_ERR_UNSUPPORTED, "lu: scratch too small");
    Cand* cand = reinterpret_cast<Cand*>(ctx->scratch);
    double* send = ctx->scratch + 2 * (size_t)grid;
    double* recv = (world > 1) ? send + stride : send;
    GSI_CUDA(cudaMemsetAsync(ctx->dflags, 0, sizeof(int), ctx->stream));
    // the blocked path needs the first l rows (the pivot rows / U) on one rank
    const int pb = (world > 1 && part_row0[1] < l) ? l : LU_PB;
    const bool own_top = (world == 1) || (ctx->rank == 0);

    const size_t smem = 2 * (size_t)l * sizeof(double);
    // experimental single-launch panels (rows all local): cooperative grid of co-resident CTAs
    LuFusedParams fp;
    int fgrid = 0;
    if (ctx->lu_fused && world == 1) {
        int occ = 0;
        GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_panel_fused_kernel, LU_THREADS, smem));
        fgrid = ctx->num_sms * (occ < 2 ? occ : 2);
        const int64_t need8 = (nloc + 7) / 8;
        if (fgrid > need8) fgrid = (int)need8;
        const size_t fneed = (size_t)2 * fgrid * 2 + (size_t)2 * fgrid * l + 2 * (size_t)l + 16;
        if (fgrid < 1 || fneed > ctx->scratch_doubles) fgrid = 0;          // fall back to the per-column path
        if (fgrid > 0) {
            fp.Y = Y->d; fp.ld = Y->ld; fp.n = nloc; fp.l = l;
            fp.cand = reinterpret_cast<Cand*>(ctx->scratch);
            fp.rows = ctx->scratch + (size_t)2 * fgrid * 2;
            fp.krows = fp.rows + (size_t)2 * fgrid * l;
            fp.flags = ctx->dflags;
            fp.R = round_up((nloc + fgrid - 1) / fgrid, 8);
        }
    }
    for (int ps = 0; ps < l; ps += pb) {
        const int pe = (ps + pb < l) ? ps + pb : l;
        if (fgrid > 0) {
            fp.ps = ps; fp.pe = pe;
            void* args[] = {&fp};
            GSI_CUDA(cudaLaunchCooperativeKernel((void*)lu_panel_fused_kernel, dim3((unsigned)fgrid), dim3(LU_THREADS),
                                                 args, smem, ctx->stream));
            count_launch(ctx);
        } else {
            lu_search_kernel<<<grid, LU_THREADS, 0, ctx->stream>>>(Y->d, Y->ld, nloc, row0, ps, cand);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
            for (int k = ps; k < pe; ++k) {
                lu_pack_kernel<<<1, LU_THREADS, 0, ctx->stream>>>(Y->d, Y->ld, nloc, row0, l, k, cand, grid, send);
                GSI_CUDA(cudaGetLastError());
                int owner_k = 0;
                if (world > 1) {
                    comm_allgather(ctx, send, recv, stride * sizeof(double));
                    for (int r = 0; r < world; ++r)
                        if (k >= part_row0[r] && k < part_row0[r + 1]) owner_k = r;
                }
                lu_eliminate_kernel<<<grid, LU_THREADS, smem, ctx->stream>>>(Y->d, Y->ld, nloc, row0, l, k, pe, recv, world,
                                                                              owner_k, cand, ctx->dflags);
                GSI_CUDA(cudaGetLastError());
                count_launch(ctx, 2);
            }
        }
        if (pe < l) {
            // U12 on the owner of the pivot rows, broadcast, then the rank-pb trailing update
            BufPtr U = make_buf(ctx, GSI_LAYOUT_TALL, pe - ps, l - pe);
            if (own_top) {
                const int ncol = l - pe;
                lu_u12_kernel<<<(ncol + 127) / 128, 128, 0, ctx->stream>>>(Y->d, Y->ld, l, ps, pe, (int64_t)ps - row0,
                                                                           U->d, U->ld);
                GSI_CUDA(cudaGetLastError());
                count_launch(ctx);
            }
            if (world > 1) comm_broadcast(ctx, U->d, (size_t)U->rows_alloc * U->ld, 0);
            int64_t i0 = (int64_t)pe - row0;
            if (i0 < 0) i0 = 0;
            if (i0 < nloc)
                tall_window_update(ctx, Y->d + i0 * Y->ld + ps, Y->ld, nloc - i0, pe - ps, U.get(),
                                   Y->d + i0 * Y->ld + pe, Y->ld, -1.0);
        }
    }
    lu_finalize_kernel<<<l, 64, 0, ctx->stream>>>(Y->d, Y->ld, nloc, row0, l);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
    int flag = 0;
    GSI_CUDA(cudaMemcpyAsync(&flag, ctx->dflags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag != 0)
        throw Error(GSI_ERR_SINGULAR, "SingularException(" + std::to_string(flag) + "): exactly zero pivot in lu");
}

}  // namespace gsi
