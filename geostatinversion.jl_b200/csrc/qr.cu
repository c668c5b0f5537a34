// Tall-skinny Householder QR with explicit thin Q:
// replaces `Matrix(qr(Y, Val(true)).Q)` (reference src/RandMatFact.jl:57-58,75-76 ->
// LAPACK dgeqp3 + dorgqr; only range(Q) matters) and the QR of B' that
// precedes the small SVD (`svd(B)`, :86).
//
// Blocked (compact-WY) algorithm with panels of QB = 16 columns, LAPACK dgeqrt/dorgqr style:
//   panel factorisation  column by column with dgeqr2/dlarfg reflectors, touching only the
//                        n x 16 panel.  Default driver (option "qr.panel" = 1): ONE cooperative
//                        launch per panel, qr_panel_kernel -- every CTA keeps its rows of the
//                        panel in SHARED MEMORY for all 16 column steps; a step is: publish the
//                        CTA's partial dot products (column k against the panel) -> one grid
//                        barrier -> every CTA reduces the partials in the same fixed order and
//                        derives the Householder scalars redundantly -> applies the reflector to
//                        its rows and accumulates the next column's dot products (4 lanes per
//                        row, warp-shuffle reductions).  Second driver ("qr.panel" = 0, the
//                        first round's scheme): per column one single-CTA kernel (scalars) and
//                        one grid pass (update), the panel served from L2;
//   T factor             G = V'V by the DMMA Gram kernel + dlarft recurrence (16 x 16);
//   trailing update      W = V'Y2 (DMMA Gram kernel, split over row chunks, deterministic
//                        two-stage reduction), W <- T'W, Y2 -= V W on the dense DMMA GEMM
//                        (tall_window_update): the trailing matrix is read/written l/16 times.
// Q is formed by the same block reflectors applied backwards (dorgqr): Q2 -= V (T (V'Q2)),
// panel columns Q1 = E - V (T V_top').  Across GPUs the R factors are all-gathered and
// re-factored redundantly (TSQR, algos.cu).
#include "common.cuh"
#include "algos.h"
#include "ptx.cuh"
#include "nb_list.h"
#include "panel_xch.cuh"

namespace gsi {

constexpr int QB = 16;               // panel width
constexpr int QP_THREADS = 256;      // panel column-step kernels
constexpr int QP_WARPS = QP_THREADS / 32;
constexpr int GR_THREADS = 512;      // Gram kernel: 16 warps = 4 row chunks x 4 column groups
constexpr int GR_RC = GR_THREADS / 32 / 4;

struct QrScal { double tau, scale, beta, pad; };

// ----------------------------------------------------------------------------- panel steps
// Accumulate, for rows i in (kdot, n):  part[jj] = sum_i Y[i, kdot] * Y[i, ps + jj]  (jj < pb).
// Shared epilogue of the panel kernels: lanes hold psum[c] for panel column (sub + 4c).
__device__ __forceinline__ void panel_store_partials(double (&psum)[4], double* __restrict__ part_cta) {
    __shared__ double s_part[QP_WARPS][QB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        double v = psum[c];
        for (int o = 4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);   // over the 8 row slots
        if (lane < 4) s_part[warp][sub + 4 * c] = v;
    }
    __syncthreads();
    if (threadIdx.x < QB) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < QP_WARPS; ++w) s += s_part[w][threadIdx.x];
        part_cta[threadIdx.x] = s;
    }
}

// dots of the first panel column with the panel columns, rows > ps
__global__ void __launch_bounds__(QP_THREADS)
qr_panel_dots_kernel(const double* __restrict__ Y, int64_t ld, int64_t n, int ps, int pe, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3, rsub = lane >> 2;
    double psum[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t wglobal = (int64_t)blockIdx.x * QP_WARPS + warp, wtotal = (int64_t)gridDim.x * QP_WARPS;
    for (int64_t i = ps + 1 + wglobal * 8 + rsub; i < n; i += wtotal * 8) {
        const double* yrow = Y + i * ld;
        const double y0 = yrow[ps];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = ps + sub + 4 * c;
            if (j < pe) psum[c] += y0 * yrow[j];
        }
    }
    panel_store_partials(psum, partial + (size_t)blockIdx.x * QB);
}

// Householder scalars of column k (dlarfg) + row-k update inside the panel.  tw[jj] = tau*w_j.
// The per-CTA partials are reduced by all 256 threads (16 slices per column, combined in a
// fixed order): deterministic, and 20 us per column faster than 16 threads walking them.
__global__ void __launch_bounds__(QP_THREADS)
qr_house_kernel(double* __restrict__ Y, int64_t ld, int ps, int pe, int k, const double* __restrict__ partial,
                 int nparts, double* __restrict__ tw, double* __restrict__ taus, QrScal* __restrict__ scal) {
    __shared__ double s_red[QP_THREADS / QB][QB];
    __shared__ double s_g[QB];
    __shared__ double s_tau, s_scale;
    {
        const int c = threadIdx.x % QB, slice = threadIdx.x / QB;          // 16 slices x 16 columns
        double s = 0.0;
        for (int b = slice; b < nparts; b += QP_THREADS / QB) s += partial[(size_t)b * QB + c];
        s_red[slice][c] = s;
    }
    __syncthreads();
    if (threadIdx.x < QB) {
        double s = 0.0;
#pragma unroll
        for (int sl = 0; sl < QP_THREADS / QB; ++sl) s += s_red[sl][threadIdx.x];
        s_g[threadIdx.x] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double alpha = Y[(int64_t)k * ld + k];
        const double xnorm2 = s_g[k - ps];
        double tau = 0.0, scale = 0.0, beta = alpha;
        if (xnorm2 > 0.0) {
            const double nrm = sqrt(alpha * alpha + xnorm2);
            beta = (alpha >= 0.0) ? -nrm : nrm;
            tau = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        Y[(int64_t)k * ld + k] = beta;
        taus[k] = tau;
        s_tau = tau; s_scale = scale;
        scal->tau = tau; scal->scale = scale; scal->beta = beta;
    }
    __syncthreads();
    const double tau = s_tau, scale = s_scale;
    const int j = k + 1 + threadIdx.x;
    if (j < pe) {
        const double ykj = Y[(int64_t)k * ld + j];
        const double w = ykj + scale * s_g[j - ps];         // v' * Y[:, j]   (v_k = 1)
        const double t = tau * w;
        Y[(int64_t)k * ld + j] = ykj - t;
        tw[j - ps] = t;
    }
}

// rows i > k: Y[i,k] <- v_i = scale*Y[i,k]; Y[i,j] -= v_i*tw[j] for panel columns j > k;
// fused: dot products of the new column k+1 with the panel columns (rows > k+1).
__global__ void __launch_bounds__(QP_THREADS)
qr_update_kernel(double* __restrict__ Y, int64_t ld, int64_t n, int ps, int pe, int k, const double* __restrict__ tw,
                 const QrScal* __restrict__ scal, double* __restrict__ partial) {
    __shared__ double s_tw[QB];
    if (threadIdx.x < QB) s_tw[threadIdx.x] = (ps + (int)threadIdx.x > k && ps + (int)threadIdx.x < pe) ? tw[threadIdx.x] : 0.0;
    __syncthreads();
    const double scale = scal->scale;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane & 3, rsub = lane >> 2;
    double psum[4] = {0.0, 0.0, 0.0, 0.0};
    const int64_t wglobal = (int64_t)blockIdx.x * QP_WARPS + warp, wtotal = (int64_t)gridDim.x * QP_WARPS;
    const int niter_guard = (int)((n - (k + 1) + wtotal * 8 - 1) / (wtotal * 8));
    for (int it = 0; it < niter_guard; ++it) {
        const int64_t i = k + 1 + (wglobal + (int64_t)it * wtotal) * 8 + rsub;
        const bool valid = i < n;
        double nv[4] = {0.0, 0.0, 0.0, 0.0};
        if (valid) {
            double* yrow = Y + i * ld;
            const double v = scale * yrow[k];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int j = ps + sub + 4 * c;
                if (j >= k && j < pe) {
                    nv[c] = (j == k) ? v : yrow[j] - v * s_tw[j - ps];
                    yrow[j] = nv[c];
                }
            }
        }
        // the new column k+1 entry of this row: panel slot (k+1-ps) -> lane sub = slot & 3, chunk slot >> 2
        const int slot = k + 1 - ps;
        double mine = 0.0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c == (slot >> 2)) mine = nv[c];
        const double ynext = __shfl_sync(0xffffffffu, mine, (lane & ~3) | (slot & 3));
        if (valid && i > k + 1) {
#pragma unroll
            for (int c = 0; c < 4; ++c) psum[c] += ynext * nv[c];
        }
    }
    panel_store_partials(psum, partial + (size_t)blockIdx.x * QB);
}


// ----------------------------------------------------------------------------- panel driver
// One cooperative launch per panel [ps, pe).  CTA b owns
// rows [b*R, (b+1)*R) of Y; its rows >= ps are active; the first `cap` owned rows live in shared
// memory (pitch QPK_PITCH), the rest
// is worked on in place.  Per column step every CTA publishes ONE record (panel_xch.cuh), then the grid barrier:
//     [ its 16 partial dot products of column k with the panel columns, row k's 16 panel entries (owner only) ].
constexpr int QPK_THREADS = 512;
constexpr int QPK_WARPS = QPK_THREADS / 32;
constexpr int QPK_PITCH = 17;        // odd: the 32 rows a warp updates (one thread per row, same column) fall into distinct bank pairs
constexpr int QPK_SLICES = QPK_THREADS / QB;      // 32 slices of the cross-CTA reduction

struct QrPanelParams {
    double* Y; int64_t ld; int64_t n; int ps, pe;
    double* recs;                   // CL = 0: [2][kPxchMaxCtas][kPxchRec]: part[16], krow[16]
    unsigned int* bar;              // CL = 0: barrier counters (panel_xch.cuh)
    unsigned int bar_base;          //         barriers of this factorisation before this launch
    double* taus;
    double* T;                      // out: the panel's T factor (QB x QB, column-major, upper triangular), written by CTA 0
    int64_t R;
    int cap;
};

// CL = 0: cooperative grid, records in global memory; CL = 1: the launch is one thread-block cluster,
// records pushed into every CTA's shared memory (panel_xch.cuh).
// Sum of ps[0..15] over the 32 lanes of a warp in 16 shuffle-adds (transposing butterfly: every stage halves the
// values a lane carries): afterwards lane L holds the warp total of column qr_col_of_lane(L) (each column in two
// lanes).  Fixed order, so deterministic.
__device__ __forceinline__ int qr_col_of_lane(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}
__device__ __forceinline__ double qr_warp_reduce16(const double (&ps)[QB], int lane) {
    double h8[8], h4[4], h2[2];
    const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4, u2 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const double send = u16 ? ps[i] : ps[i + 8], keep = u16 ? ps[i + 8] : ps[i];
        h8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const double send = u8 ? h8[i] : h8[i + 4], keep = u8 ? h8[i + 4] : h8[i];
        h4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const double send = u4 ? h4[i] : h4[i + 2], keep = u4 ? h4[i + 2] : h4[i];
        h2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const double send = u2 ? h2[0] : h2[1], keep = u2 ? h2[1] : h2[0];
    double h = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    h += __shfl_xor_sync(0xffffffffu, h, 1);
    return h;
}

template <int CL>
__global__ void __launch_bounds__(QPK_THREADS, 1) qr_panel_kernel(const QrPanelParams p) {
    extern __shared__ double sm[];                 // [cap][QPK_PITCH]
    __shared__ double s_xrec[CL ? 2 * kPxchClusterMax * kPxchRec : 1];   // CL = 1: everybody's records, by parity
    __shared__ double s_red[QPK_SLICES][QB];
    __shared__ double s_wpart[QPK_WARPS][QB];
    __shared__ double s_twv[QPK_WARPS][QB];        // CL = 1: tw of the column step, per warp
    __shared__ double s_g[QB], s_tw[QB];
    __shared__ double s_scale;
    __shared__ double s_T[QB][QB + 1], s_tau[QB], s_z[QB];   // T factor of the panel, built column by column (warp 0)
    __shared__ double s_vrow[QB];                  // row ps + c of the panel as published for column step c (kept for the T column)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & 3, rslot = tid >> 2;
    const int G = (int)gridDim.x, b = (int)blockIdx.x;       // CL = 1: the grid is one cluster
    const int pb = p.pe - p.ps;
    const int64_t r0 = (int64_t)b * p.R;
    const int64_t r1 = (r0 + p.R < p.n) ? r0 + p.R : p.n;
    const int nown = r1 > r0 ? (int)(r1 - r0) : 0;
    const int64_t ld = p.ld;
    double* Ypan = p.Y + p.ps;
    unsigned int nbar = p.bar_base;
    auto rowp = [&](int li) -> double* {
        return li < p.cap ? sm + (size_t)li * QPK_PITCH : Ypan + (r0 + li) * ld;
    };
    auto first_local = [&](int64_t grow) -> int {       // first local row with global index >= grow
        return grow > r0 ? (int)((grow - r0 < nown) ? grow - r0 : nown) : 0;
    };
    // entry j of the record CTA q published for column parity par
    auto rec_rd = [&](int par, int q, int j) -> double {
        if (CL) return s_xrec[(par * kPxchClusterMax + q) * kPxchRec + j];
        return __ldcg(p.recs + ((size_t)par * kPxchMaxCtas + q) * kPxchRec + j);
    };
    auto rec_wr = [&](int par, int j, double v) {
        if (CL) {
            cg::cluster_group cluster = cg::this_cluster();
            double* mine = s_xrec + (par * kPxchClusterMax + b) * kPxchRec + j;
            for (int d = 0; d < G; ++d) *cluster.map_shared_rank(mine, d) = v;
        } else {
            p.recs[((size_t)par * kPxchMaxCtas + b) * kPxchRec + j] = v;
        }
    };
    auto barrier = [&]() {
        if (CL) cg::this_cluster().sync();
        else panel_grid_barrier(p.bar, ++nbar, G, b);
    };
    const int lfirst = first_local(p.ps);
    const int nres = nown < p.cap ? nown : p.cap;
    for (int li = lfirst + rslot; li < nres; li += QPK_THREADS / 4) {
        const double* src = Ypan + (r0 + li) * ld;
        double* dst = sm + (size_t)li * QPK_PITCH;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = sub + 4 * c;
            if (j < pb) dst[j] = src[j];
        }
    }
    __syncthreads();
    if (CL) cg::this_cluster().sync();             // every CTA of the cluster runs before anybody pushes a record into it

    // block reduction of the lanes' psum[cc] (panel column sub + 4cc) -> my record of column step c;
    // the owner of row `col` also publishes that row
    // psum: per-thread dot products of the rows in shared memory; ovsum: lane j < 16 holds the warp's dot product of
    // column j over its rows in global memory
    auto publish = [&](int col, int c, const double (&psum)[QB], double ovsum) {
        const int par = c & 1;
        {
            const double v = qr_warp_reduce16(psum, lane);                              // over the 32 rows of the warp
            if ((lane & 1) == 0) s_wpart[warp][qr_col_of_lane(lane)] = v;
            __syncwarp();
            if (lane < QB) s_wpart[warp][lane] += ovsum;
        }
        __syncthreads();                                   // also: all rows of this CTA are up to date
        if (CL) {
            // every warp sums the warp partials (same order everywhere); warp w pushes the record into CTA w's copy
            double s = 0.0;
            if (lane < QB) {
#pragma unroll
                for (int w = 0; w < QPK_WARPS; ++w) s += s_wpart[w][lane];
            }
            if (warp < G) {
                double* dst = cg::this_cluster().map_shared_rank(s_xrec + (par * kPxchClusterMax + b) * kPxchRec, warp);
                if (lane < QB) dst[lane] = s;
                else if (lane - QB < pb && col >= r0 && col < r1) dst[lane] = rowp((int)(col - r0))[lane - QB];
            }
        } else if (tid < QB) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < QPK_WARPS; ++w) s += s_wpart[w][tid];
            rec_wr(par, tid, s);
        } else if (tid >= 32 && tid < 32 + pb) {
            if (col >= r0 && col < r1) rec_wr(par, QB + tid - 32, rowp((int)(col - r0))[tid - 32]);
        }
        barrier();                                         // every CTA's record of this column step is visible
    };

    // T factor (dlarft) without a separate pass: while the rows are updated for reflector c, the slots j < c of the
    // dot-product record (unused by the factorisation) accumulate z_c[j] = sum_{i > k} V[i, j] * v_c[i]; with the unit
    // entry of v_c this gives V[:, 0:c]' v_c, and T[0:c, c] = -tau_c * T[0:c, 0:c] * z_c.  Warp 0, lane m holding the
    // reduced slot m; cprev = c - 1 at the start of step c (and once more after the last step).  Row ps + cprev
    // was saved in s_vrow during step cprev (its record may already be overwritten: records live for one step).
    auto build_T_column = [&](int cprev, double g_lane) {
        if (lane < cprev) s_z[lane] = g_lane + s_vrow[lane];                             // + V[ps + cprev, m] * 1
        __syncwarp();
        const double tau = s_tau[cprev];
        if (lane < cprev) {
            double zz = 0.0;
            for (int m = lane; m < cprev; ++m) zz += s_T[lane][m] * s_z[m];
            s_T[lane][cprev] = -tau * zz;
        } else if (lane == cprev) {
            s_T[cprev][cprev] = tau;
        }
        __syncwarp();
    };

    // ---- dots of the first panel column with the panel columns, rows > ps (one thread per row in shared memory,
    //      16 lanes per row for the rows in global memory)
    {
        double psum[QB];
#pragma unroll
        for (int j = 0; j < QB; ++j) psum[j] = 0.0;
        const int lstart = first_local((int64_t)p.ps + 1);
        for (int li = lstart + tid; li < nres; li += QPK_THREADS) {
            const double* row = sm + (size_t)li * QPK_PITCH;
            const double y0 = row[0];
#pragma unroll
            for (int j = 0; j < QB; ++j)
                if (j < pb) psum[j] = fma(y0, row[j], psum[j]);
        }
        double ovsum = 0.0;
        {
            const int col = lane & 15, hw = lane >> 4;
            const int ostart = nres > lstart ? nres : lstart;
            constexpr int UNR = 8;
            for (int base = ostart + 2 * warp; base < nown; base += 2 * QPK_WARPS * UNR) {     // warp-uniform trip count
                double x[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int li = base + u * 2 * QPK_WARPS + hw;
                    x[u] = (li < nown && col < pb) ? Ypan[(r0 + li) * ld + col] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const double y0 = __shfl_sync(0xffffffffu, x[u], lane & 16);
                    ovsum = fma(y0, x[u], ovsum);                   // (rows past the end and columns >= pb contribute zeros)
                }
            }
            ovsum += __shfl_xor_sync(0xffffffffu, ovsum, 16);
        }
        publish(p.ps, 0, psum, ovsum);
    }

    for (int k = p.ps; k < p.pe; ++k) {
        const int c = k - p.ps;
        const int par = c & 1;
        const int owner_k = (int)(k / p.R);                                // its record carries row k
        const bool own_k = (k >= r0 && k < r1);
        const double alpha = rec_rd(par, owner_k, QB + c);                 // (issued together with the partials)
        double scale;
        double twv[QB];                                                    // tw[j] = tau * (v' Y[:, j]), 0 for j <= c
        if (CL) {
            // ---- every WARP sums the G partials of column `lane` (CTA order) and derives the Householder scalars
            //      (dlarfg) for itself: no block barrier between the cluster barrier and the update of the rows
            double gj = 0.0;
#pragma unroll
            for (int q = 0; q < kPxchClusterMax; ++q)
                if (q < G && lane < QB) gj += rec_rd(par, q, lane);
            const double ykj = (lane < pb) ? rec_rd(par, owner_k, QB + lane) : 0.0;
            if (warp == 0) {
                if (c >= 1) build_T_column(c - 1, gj);
                if (lane < QB) s_vrow[lane] = ykj;
            }
            const double xnorm2 = __shfl_sync(0xffffffffu, gj, c);
            double tau = 0.0, beta = alpha;
            scale = 0.0;
            if (xnorm2 > 0.0) {
                const double nrm = sqrt(alpha * alpha + xnorm2);
                beta = (alpha >= 0.0) ? -nrm : nrm;
                tau = (beta - alpha) / beta;
                scale = 1.0 / (alpha - beta);
            }
            double t = 0.0;
            if (lane > c && lane < pb) t = tau * (ykj + scale * gj);      // v' * Y[:, j]   (v_k = 1)
            if (warp == 0 && own_k) {                                      // row k of R; nobody reads it again in this panel
                if (lane > c && lane < pb) rowp((int)(k - r0))[lane] = ykj - t;
                else if (lane == c) rowp((int)(k - r0))[c] = beta;
            }
            if (tid == 0) {
                s_tau[c] = tau;
                if (b == 0) p.taus[k] = tau;
            }
            if (lane < QB) s_twv[warp][lane] = t;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < QB; ++j) twv[j] = s_twv[warp][j];
            __syncwarp();
        } else {
            const double ykj_mine = (tid < pb) ? rec_rd(par, owner_k, QB + tid) : 0.0;
            // ---- every CTA reduces the G partials in the same fixed order (all loads in flight before the first add)
            {
                const int j = tid % QB, slice = tid / QB;
                double pv[kPxchMaxCtas / QPK_SLICES];
#pragma unroll
                for (int it = 0; it < kPxchMaxCtas / QPK_SLICES; ++it) {
                    const int q = slice + it * QPK_SLICES;
                    pv[it] = q < G ? rec_rd(par, q, j) : 0.0;
                }
                double s = 0.0;
#pragma unroll
                for (int it = 0; it < kPxchMaxCtas / QPK_SLICES; ++it)
                    if (slice + it * QPK_SLICES < G) s += pv[it];
                s_red[slice][j] = s;
            }
            __syncthreads();
            if (tid < QB) {
                double s = 0.0;
#pragma unroll
                for (int sl = 0; sl < QPK_SLICES; ++sl) s += s_red[sl][tid];
                s_g[tid] = s;
            }
            __syncthreads();
            if (warp == 0) {
                if (c >= 1) build_T_column(c - 1, lane < QB ? s_g[lane] : 0.0);
                if (lane < QB) s_vrow[lane] = ykj_mine;      // (tid < pb holds row k's entry; 0 beyond)
            }
            // ---- Householder scalars (dlarfg), redundantly in every CTA; tw[j] = tau * (v' Y[:, j])
            {
                const double xnorm2 = s_g[c];
                double tau = 0.0, sc = 0.0, beta = alpha;
                if (xnorm2 > 0.0) {
                    const double nrm = sqrt(alpha * alpha + xnorm2);
                    beta = (alpha >= 0.0) ? -nrm : nrm;
                    tau = (beta - alpha) / beta;
                    sc = 1.0 / (alpha - beta);
                }
                if (tid == 0) {
                    s_scale = sc;
                    s_tau[c] = tau;
                    if (b == 0) p.taus[k] = tau;
                }
                if (tid < pb) {
                    const int j = tid;
                    double t = 0.0;
                    if (j > c) {
                        const double ykj = ykj_mine;
                        const double w = ykj + sc * s_g[j];            // v' * Y[:, j]   (v_k = 1)
                        t = tau * w;
                        if (own_k) rowp((int)(k - r0))[j] = ykj - t;
                    } else if (j == c && own_k) {
                        rowp((int)(k - r0))[c] = beta;
                    }
                    s_tw[j] = t;
                }
            }
            __syncthreads();
            scale = s_scale;
#pragma unroll
            for (int j = 0; j < QB; ++j) twv[j] = (j < pb) ? s_tw[j] : 0.0;
        }
        auto twv_of = [&](int j) -> double { return CL ? s_twv[warp][j] : s_tw[j]; };   // tw[j] by a run-time column index
        // ---- rows i > k: v_i = scale * Y[i,k]; Y[i,j] -= v_i * tw[j]; dots of column k+1 (rows > k+1).
        //      ONE THREAD PER ROW: the 16 panel columns of a row, the new column-(k+1) entry and its 16 products stay
        //      in one thread (no shuffles, ~70 instructions per row instead of ~100 per quarter row).
        double psum[QB];
#pragma unroll
        for (int j = 0; j < QB; ++j) psum[j] = 0.0;
        const int lstart = first_local((int64_t)k + 1);
        const int li_top = (k + 1 >= r0 && k + 1 < r1) ? (int)(k + 1 - r0) : -1;     // row k+1 itself takes no part in the dots
        auto update_row = [&](double* row, int li) {
            double nv[QB];
            const double v = scale * row[c];
            row[c] = v;
            double ynext = 0.0;
#pragma unroll
            for (int j = 0; j < QB; ++j) {
                nv[j] = 0.0;
                if (j < c) nv[j] = row[j];                  // reflector entries of the earlier columns (T factor)
                if (j == c) nv[j] = v;
                if (j > c && j < pb) {                      // (uniform over the block)
                    nv[j] = row[j] - v * twv[j];
                    row[j] = nv[j];
                    if (j == c + 1) ynext = nv[j];
                }
            }
            const double ydot = (li != li_top) ? ynext : 0.0;
#pragma unroll
            for (int j = 0; j < QB; ++j) psum[j] = fma(j < c ? v : ydot, nv[j], psum[j]);
        };
        for (int li = lstart + tid; li < nres; li += QPK_THREADS)            // rows resident in shared memory
            update_row(sm + (size_t)li * QPK_PITCH, li);
        // rows beyond the shared-memory capacity, in place in global memory: 16 LANES PER ROW, lane = column (one
        // coalesced 128-byte access per row); the lane accumulates the dot product of its own column
        double ovsum = 0.0;
        {
            const int col = lane & 15, hw = lane >> 4;
            const bool upd = col > c && col < pb;
            const double twj = upd ? twv_of(col) : 0.0;
            const int ostart = nres > lstart ? nres : lstart;
            constexpr int UNR = 8;                          // row pairs in flight per warp (the loads are the latency)
            for (int base = ostart + 2 * warp; base < nown; base += 2 * QPK_WARPS * UNR) {     // warp-uniform trip count
                double x[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int li = base + u * 2 * QPK_WARPS + hw;
                    x[u] = (li < nown && col < pb) ? Ypan[(r0 + li) * ld + col] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int li = base + u * 2 * QPK_WARPS + hw;
                    const bool valid = li < nown;
                    const double v = scale * __shfl_sync(0xffffffffu, x[u], (lane & 16) | c);
                    double nvj = 0.0;
                    if (col == c) nvj = v;
                    if (upd) nvj = x[u] - v * twj;
                    if (valid && (upd || col == c)) Ypan[(r0 + li) * ld + col] = nvj;
                    const double ynext = __shfl_sync(0xffffffffu, nvj, (lane & 16) | ((c + 1) & 15));
                    if (valid && col < c) ovsum = fma(v, x[u], ovsum);                                  // T factor: V[i, col] * v_i
                    else if (valid && li != li_top && c + 1 < pb) ovsum = fma(ynext, nvj, ovsum);
                }
            }
            ovsum += __shfl_xor_sync(0xffffffffu, ovsum, 16);          // both half-warps: lane j < 16 holds column j
        }
        publish(k + 1 < p.pe ? k + 1 : -1, c + 1, psum, ovsum);      // (after the last column: only the T-factor sums)
    }
    // ---- last column of T, then the factor goes to global memory (CTA 0)
    {
        const int par = pb & 1;
        double gl = 0.0;
        if (CL) {
#pragma unroll
            for (int q = 0; q < kPxchClusterMax; ++q)
                if (q < G && lane < QB) gl += rec_rd(par, q, lane);
        } else {
            const int j = tid % QB, slice = tid / QB;
            double sacc = 0.0;
            for (int q = slice; q < G; q += QPK_SLICES) sacc += rec_rd(par, q, j);
            s_red[slice][j] = sacc;
            __syncthreads();
            if (lane < QB) {
#pragma unroll
                for (int sl = 0; sl < QPK_SLICES; ++sl) gl += s_red[sl][lane];
            }
        }
        if (warp == 0) {
            build_T_column(pb - 1, gl);
            if (b == 0) {
                for (int idx = lane; idx < QB * QB; idx += 32) {
                    const int a = idx % QB, cc = idx / QB;
                    p.T[cc * QB + a] = (a <= cc && cc < pb) ? s_T[a][cc] : 0.0;
                }
            }
        }
    }
    __syncthreads();
    for (int li = lfirst + rslot; li < nres; li += QPK_THREADS / 4) {
        double* dst = Ypan + (r0 + li) * ld;
        const double* src = sm + (size_t)li * QPK_PITCH;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = sub + 4 * c;
            if (j < pb) dst[j] = src[j];
        }
    }
}

// ----------------------------------------------------------------------------- Gram kernel
// partial[cta][a][j] = sum over this CTA's rows of P1[i, a] * P2[i, j],  a < 16, j < 8*NB.
// P1 / P2 are windows of TALL buffers (pointer + pitch); columns >= c1 / >= c2 read as zero.
// A streaming kernel (2 flops per byte of P2 at 16 panel columns): 16 warps = 4 row chunks x 4
// column groups per CTA, and every warp issues the loads of 8-16 rows (2-4 k-steps) before the
// first MMA of the group, so that 16-30 independent loads per lane are in flight.
template <int NB>
__global__ void __launch_bounds__(GR_THREADS)
gram_kernel(const double* __restrict__ P1, int64_t ld1, int c1, const double* __restrict__ P2, int64_t ld2, int c2,
            int64_t rows, double* __restrict__ partial) {
    extern __shared__ double s_tile[];                  // [16][8*NB]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    constexpr int lp = 8 * NB;
    constexpr int NBW = (NB + 3) / 4;                   // n-blocks per warp (4 column groups)
    for (int i = threadIdx.x; i < 16 * lp; i += GR_THREADS) s_tile[i] = 0.0;
    // a row chunk is contiguous, a multiple of 16 rows
    const int cg = warp & 3, rc = warp >> 2;
    const int nb0 = cg * NBW;
    const int64_t nchunks = (int64_t)gridDim.x * GR_RC;
    int64_t chunk = (rows + nchunks - 1) / nchunks;
    chunk = (chunk + 15) / 16 * 16;
    const int64_t r0 = ((int64_t)blockIdx.x * GR_RC + rc) * chunk;
    int64_t r1 = r0 + chunk;
    if (r1 > rows) r1 = rows;
    double acc[2][NBW][2];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0.0; acc[h][nb][1] = 0.0; }
    const bool a0ok = g < c1, a1ok = 8 + g < c1;
    bool bok[NBW];
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb) bok[nb] = (nb0 + nb < NB) && ((nb0 + nb) * 8 + g < c2);
    constexpr int KS = NBW >= 8 ? 1 : (NBW >= 5 ? 2 : 4);               // k-steps (of 4 rows) loaded ahead of their MMAs: register budget
    for (int64_t i0 = r0; i0 < r1; i0 += 4 * KS) {
        double a0[KS], a1[KS], bb[KS][NBW];
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int64_t i = i0 + 4 * ks + t;
            const bool rok = i < r1;
            a0[ks] = (rok && a0ok) ? P1[i * ld1 + g] : 0.0;
            a1[ks] = (rok && a1ok) ? P1[i * ld1 + 8 + g] : 0.0;
            const double* p2 = P2 + i * ld2 + nb0 * 8 + g;
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) bb[ks][nb] = (rok && bok[nb]) ? p2[nb * 8] : 0.0;
        }
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {          // n-blocks beyond NB multiply zeros (never stored)
                dmma884(acc[0][nb][0], acc[0][nb][1], a0[ks], bb[ks][nb]);
                dmma884(acc[1][nb][0], acc[1][nb][1], a1[ks], bb[ks][nb]);
            }
        }
    }
    __syncthreads();
    // deterministic accumulation: column groups are disjoint, the row chunks add in order
    for (int w = 0; w < GR_RC; ++w) {
        if (rc == w) {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    if (nb0 + nb < NB) {
                        s_tile[(h * 8 + g) * lp + (nb0 + nb) * 8 + 2 * t] += acc[h][nb][0];
                        s_tile[(h * 8 + g) * lp + (nb0 + nb) * 8 + 2 * t + 1] += acc[h][nb][1];
                    }
                }
        }
        __syncthreads();
    }
    double* out = partial + (size_t)blockIdx.x * 16 * lp;
    for (int i = threadIdx.x; i < 16 * lp; i += GR_THREADS) out[i] = s_tile[i];
}

template <int NB>
static void launch_gram(gsi_ctx* ctx, int grid, const double* P1, int64_t ld1, int c1, const double* P2, int64_t ld2,
                        int c2, int64_t rows, double* partial) {
    const size_t smem = (size_t)16 * 8 * NB * sizeof(double);
    gram_kernel<NB><<<grid, GR_THREADS, smem, ctx->stream>>>(P1, ld1, c1, P2, ld2, c2, rows, partial);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

// returns the number of partial tiles written (grid) and the tile pitch lp
static void gram(gsi_ctx* ctx, const double* P1, int64_t ld1, int c1, const double* P2, int64_t ld2, int c2,
                 int64_t rows, double* partial, int& nparts, int& lp) {
    const int nb = nb_for_cols(c2);
    GSI_REQUIRE(nb > 0, GSI_ERR_UNSUPPORTED, "gram: too many columns");
    lp = 8 * nb;
    int64_t grid = ctx->num_sms;
    const int64_t need = (rows + GR_RC * 128 - 1) / (GR_RC * 128);
    if (grid > need) grid = need > 0 ? need : 1;
    nparts = (int)grid;
    switch (nb) {
#define GSI_CASE(N) case N: launch_gram<N>(ctx, nparts, P1, ld1, c1, P2, ld2, c2, rows, partial); break;
        GSI_NB_LIST(GSI_CASE)
#undef GSI_CASE
        default: throw Error(GSI_ERR_UNSUPPORTED, "gram: unsupported column-block count");
    }
}

// ----------------------------------------------------------------------------- small block kernels
// V_top[r][c] of the panel [ps, pe): unit lower triangular view of Y[ps+r, ps+c]
__device__ __forceinline__ double vtop(const double* __restrict__ Y, int64_t ld, int ps, int r, int c) {
    return r > c ? Y[(int64_t)(ps + r) * ld + ps + c] : (r == c ? 1.0 : 0.0);
}

// T factor of the panel (dlarft, forward columnwise): G = V'V from the Gram partials (rows >= pe)
// plus the unit-lower top block.  T is pb x pb upper triangular, column-major with ld QB.
__global__ void qr_tbuild_kernel(const double* __restrict__ Y, int64_t ld, int ps, int pe,
                                 const double* __restrict__ gpart, int nparts, int glp,
                                 const double* __restrict__ taus, double* __restrict__ T) {
    __shared__ double G[QB][QB], Ts[QB][QB];
    const int pb = pe - ps;
    const int a = threadIdx.x / QB, c = threadIdx.x % QB;     // 256 threads
    double s = 0.0;
    if (a < pb && c < pb && a < c) {
        const double* gp = gpart + (size_t)a * glp + c;
#pragma unroll 8
        for (int b = 0; b < nparts; ++b) s += gp[(size_t)b * 16 * glp];
        for (int r = c; r < pb; ++r) s += vtop(Y, ld, ps, r, a) * vtop(Y, ld, ps, r, c);
    }
    G[a][c] = s;
    Ts[a][c] = 0.0;
    __syncthreads();
    // dlarft recurrence, column by column; the rows of a column are independent (thread r: same
    // summation order as a serial walk)
    for (int cc = 0; cc < pb; ++cc) {
        const int r = threadIdx.x;
        if (r <= cc) {
            const double tau = taus[ps + cc];
            // T[0:cc, cc] = -tau * T[0:cc, 0:cc] * G[0:cc, cc]
            if (r < cc) {
                double z = 0.0;
                for (int m = r; m < cc; ++m) z += Ts[r][m] * G[m][cc];
                Ts[r][cc] = -tau * z;
            } else {
                Ts[cc][cc] = tau;
            }
        }
        __syncthreads();
    }
    T[c * QB + a] = Ts[a][c];
}

// A CTA of 256 threads handles 16 trailing columns (window columns jj0 .. jj0+15): W = V'Y2 (Gram
// partials over rows >= pe, summed by all threads: thread (c, jj) walks the nparts partials of its
// own entry, + the top block), W2 = op(T) W (transT = 1: T'W, factorisation; 0: T W, forming Q),
// top rows Y2[ps+r, j] -= sum_c V_top[r][c] W2[c], and W2 -> TALL buffer for the rank-16 update.
__global__ void __launch_bounds__(QB * QB)
qr_wt_kernel(double* __restrict__ Y, int64_t ld, int ps, int pe, int j0, int ncols,
             const double* __restrict__ wpart, int nparts, int wlp, const double* __restrict__ T,
             int transT, double* __restrict__ W2, int64_t ldw2) {
    __shared__ double Ts[QB][QB], Vt[QB][QB], y2[QB][QB + 1], w[QB][QB + 1], w2[QB][QB + 1];
    const int pb = pe - ps;
    const int a = threadIdx.x / QB, jl = threadIdx.x % QB;          // a: panel index (row r / column c), jl: local column
    const int jj = blockIdx.x * QB + jl;
    const bool colok = jj < ncols;
    const int j = j0 + jj;
    Ts[a][jl] = (a < pb && jl < pb) ? T[jl * QB + a] : 0.0;         // Ts[r][c] = T(r, c)
    Vt[a][jl] = (a < pb && jl < pb) ? vtop(Y, ld, ps, a, jl) : 0.0;
    y2[a][jl] = (a < pb && colok) ? Y[(int64_t)(ps + a) * ld + j] : 0.0;
    __syncthreads();
    {
        double s = 0.0;
        if (a < pb && colok) {
            const double* wp = wpart + (size_t)a * wlp + jj;
#pragma unroll 8
            for (int b = 0; b < nparts; ++b) s += wp[(size_t)b * 16 * wlp];
#pragma unroll
            for (int r = 0; r < QB; ++r) s += Vt[r][a] * y2[r][jl];
        }
        w[a][jl] = s;
    }
    __syncthreads();
    {
        double s = 0.0;
#pragma unroll
        for (int m = 0; m < QB; ++m) s += (transT ? Ts[m][a] : Ts[a][m]) * w[m][jl];
        w2[a][jl] = s;
    }
    __syncthreads();
    if (a < pb && colok) {
        double s = y2[a][jl];
#pragma unroll
        for (int c = 0; c < QB; ++c) s -= Vt[a][c] * w2[c][jl];
        Y[(int64_t)(ps + a) * ld + j] = s;
        W2[(int64_t)a * ldw2 + jj] = w2[a][jl];
    }
}

// M = T * V_top' (pb x pb, row-major [a][c]) for the panel columns of Q
__global__ void org_m_kernel(const double* __restrict__ Y, int64_t ld, int ps, int pe, const double* __restrict__ T,
                             double* __restrict__ Mout) {
    const int pb = pe - ps;
    const int a = threadIdx.x / QB, c = threadIdx.x % QB;
    double s = 0.0;
    if (a < pb && c < pb)
        for (int m = a; m < pb; ++m) s += T[m * QB + a] * vtop(Y, ld, ps, c, m);       // T[a][m] * V_top[c][m]
    Mout[a * QB + c] = s;
}

// Panel columns of Q: Q[:, ps:pe] = E - V M.  One thread per row i >= ps (each thread reads and
// rewrites only its own row).
__global__ void org_panel_kernel(double* __restrict__ Y, int64_t ld, int64_t n, int ps, int pe,
                                 const double* __restrict__ Min) {
    __shared__ double M[QB][QB];
    const int pb = pe - ps;
    if (threadIdx.x < QB * QB) M[threadIdx.x / QB][threadIdx.x % QB] = Min[threadIdx.x];
    __syncthreads();
    const int64_t i = ps + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* yrow = Y + i * ld + ps;
    double v[QB];
    const int r = (int)(i - ps);
#pragma unroll
    for (int a = 0; a < QB; ++a) {
        double x = 0.0;
        if (a < pb) {
            if (i >= pe) x = yrow[a];
            else x = (r > a) ? yrow[a] : (r == a ? 1.0 : 0.0);
        }
        v[a] = x;
    }
#pragma unroll
    for (int c = 0; c < QB; ++c) {
        if (c < pb) {
            double s = (i < pe && r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int a = 0; a < QB; ++a) s -= v[a] * M[a][c];
            yrow[c] = s;
        }
    }
}

__global__ void extract_R_kernel(const double* __restrict__ Y, int64_t ld, int l, double* __restrict__ R) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l * l) return;
    const int r = idx % l, c = idx / l;
    R[(size_t)c * l + r] = (r <= c) ? Y[(int64_t)r * ld + c] : 0.0;
}

// the R entries (on and above the diagonal) are not part of Q
__global__ void zero_upper_kernel(double* __restrict__ Y, int64_t ld, int l) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l * l) return;
    const int r = idx % l, c = idx / l;
    if (r <= c) Y[(int64_t)r * ld + c] = 0.0;
}

void qr_thinQ_inplace(gsi_ctx* ctx, gsi_buf* Y, double* Rdev) {
    GSI_REQUIRE(Y->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "qr: TALL buffer required");
    const int l = (int)Y->cols;
    const int64_t n = Y->rows;
    GSI_REQUIRE(n >= l, GSI_ERR_UNSUPPORTED, "qr: fewer (local) rows than columns is not supported");
    cudaStream_t st = ctx->stream;
    int grid = ctx->num_sms * 2;
    const int64_t need = (n + QP_WARPS * 8 - 1) / (QP_WARPS * 8);
    if (grid > need) grid = (int)(need > 0 ? need : 1);
    const int npanels = (l + QB - 1) / QB;
    const int lpmax = 8 * nb_for_cols(l < kMaxCols ? l : kMaxCols);     // Gram partial tiles cover <= 256 columns at a time
    // scratch: partial[grid*QB] | tw[QB] | scal | taus[l] | T[npanels*QB*QB] | gpart[num_sms*16*lpmax]
    double* partial = ctx->scratch;
    double* tw = partial + (size_t)grid * QB;
    QrScal* scal = reinterpret_cast<QrScal*>(tw + QB);
    double* taus = tw + QB + 4;
    double* Tall = taus + l;
    double* Mscr = Tall + (size_t)npanels * QB * QB;
    double* gpart = Mscr + QB * QB;
    const size_t need_doubles = (size_t)(gpart - ctx->scratch) + (size_t)ctx->num_sms * 16 * lpmax;
    GSI_REQUIRE(need_doubles <= ctx->scratch_doubles, GSI_ERR_UNSUPPORTED, "qr: scratch too small");

    // panel driver set-up.  Short iterates (qr.panel = 1): the launch is ONE thread-block cluster of up to
    // 16 CTAs; otherwise (or qr.panel = 2) a cooperative grid of co-resident CTAs, at most one per SM.
    QrPanelParams pp;
    int pgrid = 0, pmode = 0;                 // pmode 1: cluster transport
    size_t psmem = 0;
    auto panel_setup = [&](int mode) {
        pgrid = 0; pmode = mode; psmem = 0;
        const void* kfn = mode ? (const void*)qr_panel_kernel<1> : (const void*)qr_panel_kernel<0>;
        cudaFuncAttributes fa;
        GSI_CUDA(cudaFuncGetAttributes(&fa, kfn));
        const int cap_max = (int)((232448 - fa.sharedSizeBytes) / (QPK_PITCH * sizeof(double)));   // 227 KB per CTA, minus the static part
        const int ctas_max = mode ? kPxchClusterMax : ctx->num_sms;
        int64_t R = round_up((n + ctas_max - 1) / ctas_max, 8);
        if (R < 256) R = 256;
        if (mode && R > cap_max + 256) return;                          // too long for one cluster: cooperative grid
        const int g = (int)((n + R - 1) / R);
        int cap = (int)R;
        if (cap > cap_max) cap = cap_max;
        psmem = (size_t)cap * QPK_PITCH * sizeof(double);
        GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        if (mode) {
            if (g > 8) GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        } else {
            int occ = 0;
            GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qr_panel_kernel<0>, QPK_THREADS, psmem));
            if (occ < 1 || g > ctx->num_sms * occ || g > kPxchMaxCtas) return;   // per-column driver instead
        }
        pgrid = g;
        pp.Y = Y->d; pp.ld = Y->ld; pp.n = n;
        pp.recs = ctx->pxch;
        pp.bar = ctx->pbar; pp.bar_base = 0;
        pp.taus = taus;
        pp.T = nullptr;
        pp.R = R; pp.cap = cap;
    };
    if (ctx->qr_panel == 1) panel_setup(1);
    if (ctx->qr_panel && pgrid == 0) panel_setup(0);
    if (pgrid > 0 && pmode == 0) GSI_CUDA(cudaMemsetAsync(ctx->pbar, 0, kPbarBytes, st));
    auto panel_launch = [&]() -> cudaError_t {
        void* args[] = {&pp};
        if (pmode == 0)
            return cudaLaunchCooperativeKernel((void*)qr_panel_kernel<0>, dim3((unsigned)pgrid), dim3(QPK_THREADS), args,
                                               psmem, st);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pgrid);
        cfg.blockDim = dim3(QPK_THREADS);
        cfg.dynamicSmemBytes = psmem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)pgrid;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelExC(&cfg, (const void*)qr_panel_kernel<1>, args);
    };
    // ---------------- factorisation, panel by panel
    for (int ps = 0, pi = 0; ps < l; ps += QB, ++pi) {
        const int pe = (ps + QB < l) ? ps + QB : l;
        double* T = Tall + (size_t)pi * QB * QB;
        bool done = false;
        if (pgrid > 0) {
            pp.ps = ps; pp.pe = pe; pp.T = T;
            cudaError_t e = panel_launch();
            if (e != cudaSuccess && pmode == 1) {
                cudaGetLastError();            // the cluster could not be scheduled: cooperative grid from here on
                panel_setup(0);
                if (pgrid > 0) {
                    GSI_CUDA(cudaMemsetAsync(ctx->pbar, 0, kPbarBytes, st));
                    pp.ps = ps; pp.pe = pe; pp.T = T;
                    e = panel_launch();
                }
            }
            if (pgrid > 0 && e == cudaSuccess) {
                count_launch(ctx);
                pp.bar_base += (unsigned int)(pe - ps) + 1; // one barrier per column step + the T-factor exchange
                done = true;
            } else {
                cudaGetLastError();            // the grid could not be made co-resident: per-column driver
                pgrid = 0;
            }
        }
        if (!done) {
            qr_panel_dots_kernel<<<grid, QP_THREADS, 0, st>>>(Y->d, Y->ld, n, ps, pe, partial);
            for (int k = ps; k < pe; ++k) {
                qr_house_kernel<<<1, QP_THREADS, 0, st>>>(Y->d, Y->ld, ps, pe, k, partial, grid, tw, taus, scal);
                qr_update_kernel<<<grid, QP_THREADS, 0, st>>>(Y->d, Y->ld, n, ps, pe, k, tw, scal, partial);
            }
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx, 1 + 2 * (pe - ps));
        }
        const int64_t rows_below = n - pe;
        if (!done) {
            // T factor: G = V'V (rows >= pe through the Gram kernel, top block inside qr_tbuild); the panel kernel
            // delivers T itself
            int nparts = 0, glp = 0;
            gram(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, pe - ps, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, pe - ps,
                 rows_below, gpart, nparts, glp);
            qr_tbuild_kernel<<<1, QB * QB, 0, st>>>(Y->d, Y->ld, ps, pe, gpart, nparts, glp, taus, T);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
        }
        if (pe < l) {
            // trailing update: Y2 <- (I - V T' V') Y2, in chunks of at most 256 trailing columns (wide iterates)
            for (int c0 = pe; c0 < l; c0 += kMaxCols) {
                const int ncols = (l - c0 < kMaxCols) ? l - c0 : kMaxCols;
                int wparts = 0, wlp = 0;
                gram(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, pe - ps, Y->d + (int64_t)pe * Y->ld + c0, Y->ld, ncols,
                     rows_below, gpart, wparts, wlp);
                BufPtr W2 = make_buf(ctx, GSI_LAYOUT_TALL, pe - ps, ncols);
                qr_wt_kernel<<<(ncols + QB - 1) / QB, QB * QB, 0, st>>>(Y->d, Y->ld, ps, pe, c0, ncols, gpart, wparts, wlp, T, 1,
                                                                       W2->d, W2->ld);
                GSI_CUDA(cudaGetLastError());
                count_launch(ctx);
                tall_window_update(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, rows_below, pe - ps, W2.get(),
                                   Y->d + (int64_t)pe * Y->ld + c0, Y->ld, -1.0);
            }
        }
    }
    if (Rdev) {
        extract_R_kernel<<<(l * l + 255) / 256, 256, 0, st>>>(Y->d, Y->ld, l, Rdev);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
    // ---------------- explicit Q, panels backwards (dorgqr)
    zero_upper_kernel<<<(l * l + 255) / 256, 256, 0, st>>>(Y->d, Y->ld, l);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
    for (int pi = npanels - 1; pi >= 0; --pi) {
        const int ps = pi * QB;
        const int pe = (ps + QB < l) ? ps + QB : l;
        const double* T = Tall + (size_t)pi * QB * QB;
        const int64_t rows_below = n - pe;
        for (int c0 = pe; c0 < l; c0 += kMaxCols) {
            const int ncols = (l - c0 < kMaxCols) ? l - c0 : kMaxCols;
            int wparts = 0, wlp = 0;
            gram(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, pe - ps, Y->d + (int64_t)pe * Y->ld + c0, Y->ld, ncols,
                 rows_below, gpart, wparts, wlp);
            BufPtr W2 = make_buf(ctx, GSI_LAYOUT_TALL, pe - ps, ncols);
            qr_wt_kernel<<<(ncols + QB - 1) / QB, QB * QB, 0, st>>>(Y->d, Y->ld, ps, pe, c0, ncols, gpart, wparts, wlp, T, 0,
                                                                   W2->d, W2->ld);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
            tall_window_update(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, rows_below, pe - ps, W2.get(),
                               Y->d + (int64_t)pe * Y->ld + c0, Y->ld, -1.0);
        }
        const int64_t prow = n - ps;
        org_m_kernel<<<1, QB * QB, 0, st>>>(Y->d, Y->ld, ps, pe, T, Mscr);
        org_panel_kernel<<<(unsigned)((prow + 255) / 256), 256, 0, st>>>(Y->d, Y->ld, n, ps, pe, Mscr);
        count_launch(ctx);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
}

}  // namespace gsi
