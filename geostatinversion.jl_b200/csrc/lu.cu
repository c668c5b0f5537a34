// Reference-faithful normaliser:  `F = lu(Y); Q = F.L`
// (reference src/RandMatFact.jl:60-61,68-69,72-73 -> LAPACK dgetrf).
//
// Gaussian elimination with partial pivoting on a tall n x l TALL buffer, in place,
// with LAPACK's pivot rule (idamax: FIRST row of maximal |value|), physical row
// interchanges and multiplication by the reciprocal pivot (dgetf2/dgetrf2).  The result
// is the unit-lower-trapezoidal factor in LAPACK's *permuted* row order -- the reference
// never un-permutes it.
//
// Blocked right-looking algorithm (panel width LU_PB = 16): inside a panel only the panel's
// own columns are eliminated; after the panel, U12 = L11^{-1} A12 is formed by one small
// kernel and the trailing matrix receives ONE rank-16 update on the FP64 tensor cores
// (tall_window_update -> dense DMMA GEMM), i.e. it is read and written l/16 times.
//
// Panel factorisation, default driver (option "lu.panel" = 1): ONE cooperative launch per
// panel, lu_panel_kernel.  Every CTA keeps its rows of the n x 16 panel in SHARED MEMORY for
// the whole panel (rows beyond the shared-memory capacity stay in global memory / L2), so a
// column step is: block arg-max -> publish ONE record (|v|, row, the candidate's 16 panel entries,
// row k's entries) -> ONE grid barrier -> every CTA reduces the G candidates identically -> row
// interchange inside the panel (each row by its owner) -> rank-1 update of its rows + the next
// column's arg-max.
// The interchanges of the columns outside the panel are applied afterwards in one pass
// (LAPACK's dlaswp order), by one CTA, while the others write their panel rows back.
// l = 210: 14 launches with 16 grid barriers each instead of 420 launches + a host sync.
//
// Second driver (option "lu.panel" = 0, the first round's default, kept as the independent
// implementation the panel kernel is tested against bit for bit): one column = two launches
//   lu_pack      (1 CTA)  reduce the per-CTA pivot candidates of column k, export the
//                         candidate row and row k                      -> exchange buffer
//   lu_eliminate (grid)   pick the pivot, move rows k <-> p from the exchanged copies, scale
//                         column k, rank-1 update of the panel columns and -- fused -- the
//                         arg-max search of column k+1.
//
// Multi-GPU: the iterate is gathered BEFORE the normalisation (the next product needs all of
// it on every rank anyway) and every rank factors all rows redundantly (algos.cu): the kernels
// are deterministic, so the factor is identical on every rank and identical to the single-GPU
// result, and no per-column pivot exchange over NCCL exists any more (round 1: 840 small
// all-gathers per LU, 21 of the 31 ms an LU took at 8 GPUs).
//
// A NaN in the pivot column wins the arg-max (|NaN| is treated as +inf), so non-finite input
// propagates into L the way it does through LAPACK instead of leaving stale pivot choices.
// An exactly zero pivot sets the SingularException flag (checked by the caller at its next
// synchronisation point: lu_check_singular).
#include "common.cuh"
#include "algos.h"
#include "panel_xch.cuh"
#include <cmath>

namespace gsi {

constexpr int LU_PB = 16;            // panel width
constexpr int LU_THREADS = 256;
constexpr int LU_WARPS = LU_THREADS / 32;
constexpr int LP_THREADS = 512;      // panel kernel: 16 warps, 4 lanes per row -> 128 rows per pass
constexpr int LP_WARPS = LP_THREADS / 32;
constexpr int LP_PITCH = 17;         // shared-memory row pitch (doubles), odd: the 32 rows a warp eliminates (one thread per row, same
                                     // column) fall into distinct bank pairs

struct Cand { double val; double idx; };   // idx = global row index (exact in a double)

__device__ __forceinline__ bool cand_better(double v, double i, double bv, double bi) {
    return (v > bv) || (v == bv && i < bi);
}
// |x| as a pivot candidate: NaN counts as +inf so that it is chosen (and propagates)
__device__ __forceinline__ double cand_abs(double x) {
    const double a = fabs(x);
    return (a == a) ? a : INFINITY;
}

// ---------------------------------------------------------------------------- per-column driver
// column-0 search (later columns are searched inside lu_eliminate)
__global__ void lu_search_kernel(const double* __restrict__ Y, int64_t ld, int64_t nloc, int col,
                                 Cand* __restrict__ cand) {
    double bv = -1.0, bi = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nloc; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < col) continue;
        const double v = cand_abs(Y[i * ld + col]);
        if (cand_better(v, (double)i, bv, bi)) { bv = v; bi = (double)i; }
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

// xch layout: [0] = |candidate| (-1: no row >= k), [1] = row, [2 .. 2+l) candidate row, [2+l .. 2+2l) row k
__global__ void lu_pack_kernel(const double* __restrict__ Y, int64_t ld, int64_t nloc, int l, int k,
                               const Cand* __restrict__ cand, int ncand, double* __restrict__ xch) {
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
    if (threadIdx.x == 0) { xch[0] = bv; xch[1] = bi; }
    if (bv >= 0.0) {
        const int64_t li = (int64_t)bi;
        for (int j = threadIdx.x; j < l; j += LU_THREADS) xch[2 + j] = Y[li * ld + j];
    }
    if (k < nloc)
        for (int j = threadIdx.x; j < l; j += LU_THREADS) xch[2 + l + j] = Y[(int64_t)k * ld + j];
}

// Elimination of column k restricted to the panel columns (k, jend).  8 rows per warp,
// 4 lanes per row (a panel row segment is <= 15 contiguous doubles).
__global__ void __launch_bounds__(LU_THREADS)
lu_eliminate_kernel(double* __restrict__ Y, int64_t ld, int64_t nloc, int l, int k, int jend,
                    const double* __restrict__ xch, Cand* __restrict__ cand, int* __restrict__ flags) {
    extern __shared__ double sm[];
    double* prow = sm;            // pivot row (all l columns)
    double* krow = sm + l;        // previous content of row k
    __shared__ double s_best[LU_WARPS], s_bidx[LU_WARPS];
    const int64_t p = (int64_t)xch[1];
    for (int j = threadIdx.x; j < l; j += LU_THREADS) {
        prow[j] = xch[2 + j];
        krow[j] = xch[2 + l + j];
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
        for (int j = threadIdx.x; j < l; j += LU_THREADS) Y[(int64_t)k * ld + j] = prow[j];
        for (int j = threadIdx.x; j < l; j += LU_THREADS)
            if (j < k || j >= jend) Y[p * ld + j] = krow[j];
    }

    const int sub = lane & 3, rsub = lane >> 2;
    double bv = -1.0, bi = 0.0;
    const int64_t wglobal = (int64_t)blockIdx.x * LU_WARPS + warp;
    const int64_t wtotal = (int64_t)gridDim.x * LU_WARPS;
    for (int64_t i = (int64_t)k + 1 + wglobal * 8 + rsub; i < nloc; i += wtotal * 8) {
        double* yrow = Y + i * ld;
        const double* src = (i == p) ? krow : yrow;
        double m;
        if (pivot == 0.0) m = src[k];
        else m = use_recip ? src[k] * rpiv : src[k] / pivot;
        if (sub == 0) yrow[k] = m;
        for (int j = k + 1 + sub; j < jend; j += 4) {
            const double v = fma(-m, prow[j], src[j]);
            yrow[j] = v;
            if (j == k + 1) {   // sub == 0: candidate for the next column of this panel
                const double a = cand_abs(v);
                if (cand_better(a, (double)i, bv, bi)) { bv = a; bi = (double)i; }
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

// ---------------------------------------------------------------------------- panel driver
// One launch per panel [ps, pe).  CTA b owns rows [b*R, (b+1)*R) of Y (R a multiple of 8); its rows
// >= ps are the "active" ones.  The first `cap` owned rows live in shared memory (pitch LP_PITCH),
// the rest is worked on in place.
//
// Exchange (panel_xch.cuh): per column step every CTA publishes ONE record, then the barrier:
//     [ |candidate|, its row index, the candidate row's 16 panel entries, row k's 16 panel entries (owner only) ].
// CL = 0: cooperative launch of up to one CTA per SM, records in global memory, two-level counter barrier;
// CL = 1: the launch is a single thread-block cluster, records pushed into every CTA's shared memory,
//         hardware cluster barrier (short iterates).
struct LuPanelParams {
    double* Y; int64_t ld; int64_t n; int l; int ps, pe;
    double* recs;                   // CL = 0: [2][kPxchMaxCtas][kPxchRec]: val, idx, -, -, row[16], krow[16]
    unsigned int* bar;              // CL = 0: barrier counters (panel_xch.cuh)
    unsigned int bar_base;          //         barriers of this factorisation before this launch
    int* flags;
    int64_t R;
    int cap;
};

// Pivot candidates inside the panel kernel are (key, row): key = bit pattern of |value| + 1 (non-negative doubles
// order like unsigned integers; NaN counts as +inf; key 0 = no candidate), row = global row index (< 2^31).
// "better" = larger key, then smaller row -- LAPACK's idamax rule.  A warp arg-max is three redux.sync
// instructions (high word, low word, row) instead of a five-stage shuffle butterfly on two doubles.
__device__ __forceinline__ unsigned long long lp_key(double v) {
    return (unsigned long long)__double_as_longlong(cand_abs(v)) + 1ull;
}
__device__ __forceinline__ void lp_warp_argmax(unsigned long long& key, int& row) {
    const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
    const unsigned int mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned int ml = __reduce_max_sync(0xffffffffu, hi == mh ? lo : 0u);
    const bool mine = (hi == mh) && (lo == ml);
    const unsigned int mr = __reduce_min_sync(0xffffffffu, mine ? (unsigned int)row : 0xffffffffu);
    key = ((unsigned long long)mh << 32) | ml;
    row = (int)mr;
}

template <int CL>
__global__ void __launch_bounds__(LP_THREADS, 1) lu_panel_kernel(const LuPanelParams p) {
    extern __shared__ double sm[];                 // [cap][LP_PITCH]
    __shared__ double s_xrec[CL ? 2 * kPxchClusterMax * kPxchRec : 1];   // CL = 1: everybody's records, by parity
    __shared__ unsigned long long s_wkey[LP_WARPS];
    __shared__ int s_wrow[LP_WARPS];
    __shared__ double s_prow[LU_PB], s_krow[LU_PB];
    __shared__ int s_win;                          // CL = 0: pivot row index
    __shared__ int s_piv[LU_PB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, sub = lane & 3, rslot = tid >> 2;
    const int G = (int)gridDim.x, b = (int)blockIdx.x;       // CL = 1: the grid is one cluster, block rank == blockIdx.x
    const int pb = p.pe - p.ps;
    const int64_t r0 = (int64_t)b * p.R;
    const int64_t r1 = (r0 + p.R < p.n) ? r0 + p.R : p.n;
    const int nown = r1 > r0 ? (int)(r1 - r0) : 0;
    const int64_t ld = p.ld;
    double* Ypan = p.Y + p.ps;                      // panel window of the iterate
    unsigned int nbar = p.bar_base;
    // panel segment of my local row li (global row r0 + li)
    auto rowp = [&](int li) -> double* {
        return li < p.cap ? sm + (size_t)li * LP_PITCH : Ypan + (r0 + li) * ld;
    };
    // entry j of the record CTA q published for column parity par
    //   [0] key (bit pattern)  [1] row  [4 .. 20) the candidate row's panel entries  [20 .. 36) row k's entries (owner only)
    auto rec_rd = [&](int par, int q, int j) -> double {
        if (CL) return s_xrec[(par * kPxchClusterMax + q) * kPxchRec + j];
        return __ldcg(p.recs + ((size_t)par * kPxchMaxCtas + q) * kPxchRec + j);
    };
    auto barrier = [&]() {
        if (CL) cg::this_cluster().sync();
        else panel_grid_barrier(p.bar, ++nbar, G, b);
    };
    // ---- load my active rows into shared memory
    const int lfirst = (p.ps > r0) ? (int)((p.ps - r0 < nown) ? p.ps - r0 : nown) : 0;   // first active local row
    const int nres = nown < p.cap ? nown : p.cap;
    for (int li = lfirst + rslot; li < nres; li += LP_THREADS / 4) {
        const double* src = Ypan + (r0 + li) * ld;
        double* dst = sm + (size_t)li * LP_PITCH;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = sub + 4 * c;
            if (j < pb) dst[j] = src[j];
        }
    }
    __syncthreads();
    if (CL) cg::this_cluster().sync();             // every CTA of the cluster runs before anybody pushes a record into it

    // block arg-max of my (key, row) -> my record of column step c: candidate + its row, and row `col` if I own it
    auto publish = [&](int col, int c, unsigned long long key, int row) {
        const int par = c & 1;
        lp_warp_argmax(key, row);
        if (lane == 0) { s_wkey[warp] = key; s_wrow[warp] = row; }
        __syncthreads();                                   // also: all rows of this CTA are up to date
        // every warp reduces the warp candidates (identically)
        key = lane < LP_WARPS ? s_wkey[lane] : 0ull;
        row = lane < LP_WARPS ? s_wrow[lane] : 0x7fffffff;
        lp_warp_argmax(key, row);
        const bool own_col = col >= r0 && col < r1;
        if (CL) {
            // warp w pushes the record into CTA w's copy
            if (warp < G) {
                double* dst = cg::this_cluster().map_shared_rank(s_xrec + (par * kPxchClusterMax + b) * kPxchRec, warp);
                if (lane == 0) { dst[0] = __longlong_as_double((long long)key); dst[1] = (double)row; }
                if (key != 0ull && lane < pb) dst[4 + lane] = rowp((int)((int64_t)row - r0))[lane];
                if (own_col && lane >= 16 && lane < 16 + pb) dst[4 + LU_PB + lane - 16] = rowp((int)(col - r0))[lane - 16];
            }
        } else if (warp == 0) {
            double* dst = p.recs + ((size_t)par * kPxchMaxCtas + b) * kPxchRec;
            if (lane == 0) { dst[0] = __longlong_as_double((long long)key); dst[1] = (double)row; }
            if (key != 0ull && lane < pb) dst[4 + lane] = rowp((int)((int64_t)row - r0))[lane];
            if (own_col && lane >= 16 && lane < 16 + pb) dst[4 + LU_PB + lane - 16] = rowp((int)(col - r0))[lane - 16];
        }
        barrier();                                         // every CTA's record of this column step is visible
    };

    // ---- candidates of the first column of the panel
    {
        unsigned long long key = 0ull;
        int row = 0x7fffffff;
        for (int li = lfirst + tid; li < nown; li += LP_THREADS) {
            const unsigned long long kk = lp_key(rowp(li)[0]);
            if (kk > key) { key = kk; row = (int)(r0 + li); }
        }
        publish(p.ps, 0, key, row);
    }

    for (int k = p.ps; k < p.pe; ++k) {
        const int c = k - p.ps;
        const int par = c & 1;
        const int owner_k = (int)(k / p.R);
        // ---- global pivot: the G candidates reduced the same way everywhere (largest key, then smallest row).
        //      CL = 1: every warp does it for itself from the shared-memory records (no staging);
        //      CL = 0: warp 0 reads the records from L2 and stages the pivot row / row k for the block.
        int piv;
        const double* prow;                                 // the pivot row's panel entries
        const double* krow;                                 // row k's panel entries before the interchange
        if (CL) {
            unsigned long long key = lane < G ? (unsigned long long)__double_as_longlong(rec_rd(par, lane, 0)) : 0ull;
            int row = lane < G ? (int)rec_rd(par, lane, 1) : 0x7fffffff;
            const unsigned long long mykey = key;
            const int myrow = row;
            lp_warp_argmax(key, row);
            const unsigned int who = __ballot_sync(0xffffffffu, mykey == key && myrow == row && lane < G);
            krow = s_xrec + (par * kPxchClusterMax + owner_k) * kPxchRec + 4 + LU_PB;
            if (key == 0ull) { row = k; prow = krow; }      // no row left (cannot happen for n >= l): no interchange
            else prow = s_xrec + (par * kPxchClusterMax + (__ffs(who) - 1)) * kPxchRec + 4;
            piv = row;
            if (tid == 0) s_piv[c] = piv;
        } else {
            if (warp == 0) {
                unsigned long long key = 0ull;
                int row = 0x7fffffff, bw = 0;
                double kv[kPxchMaxCtas / 32], rv[kPxchMaxCtas / 32];
#pragma unroll
                for (int it = 0; it < kPxchMaxCtas / 32; ++it) {          // all loads in flight before the first compare
                    const int q = lane + 32 * it;
                    kv[it] = q < G ? rec_rd(par, q, 0) : 0.0;
                    rv[it] = q < G ? rec_rd(par, q, 1) : 0.0;
                }
#pragma unroll
                for (int it = 0; it < kPxchMaxCtas / 32; ++it) {
                    const unsigned long long kk = (lane + 32 * it < G) ? (unsigned long long)__double_as_longlong(kv[it]) : 0ull;
                    const int rr = (int)rv[it];
                    if (kk > key || (kk == key && kk != 0ull && rr < row)) { key = kk; row = rr; bw = lane + 32 * it; }
                }
                const unsigned long long mykey = key;
                const int myrow = row;
                lp_warp_argmax(key, row);
                const unsigned int who = __ballot_sync(0xffffffffu, mykey == key && myrow == row);
                bw = __shfl_sync(0xffffffffu, bw, __ffs(who) - 1);
                if (key == 0ull) row = k;                    // no row left (cannot happen for n >= l): no interchange
                if (lane < pb) {
                    s_krow[lane] = rec_rd(par, owner_k, 4 + LU_PB + lane);
                    s_prow[lane] = key != 0ull ? rec_rd(par, bw, 4 + lane) : s_krow[lane];
                }
                if (lane == 0) { s_win = row; s_piv[c] = row; }
            }
            __syncthreads();
            piv = s_win;
            prow = s_prow;
            krow = s_krow;
        }
        const double pivot = prow[c];
        double rpiv = 0.0;
        if (pivot == 0.0) {
            if (b == 0 && tid == 0) atomicCAS(&p.flags[0], 0, k + 1);   // first zero pivot (1-based)
        } else {
            rpiv = 1.0 / pivot;
        }
        // LAPACK dgetf2: multiply by the reciprocal when |pivot| >= sfmin, divide otherwise; a zero pivot leaves the column
        const int mode = pivot == 0.0 ? 2 : (fabs(pivot) >= 2.2250738585072014e-308 ? 0 : 1);
        // the pivot row's entries right of column c (0 elsewhere: those columns are left alone)
        double pr[LU_PB];
#pragma unroll
        for (int j = 0; j < LU_PB; ++j) pr[j] = (j > c && j < pb) ? prow[j] : 0.0;
        const bool has_next = c + 1 < pb;
        // ---- row interchange inside the panel, each row by its owner
        if (piv != k) {
            if (tid < pb) {
                if (k >= r0 && k < r1) rowp((int)(k - r0))[tid] = prow[tid];
            } else if (tid >= 32 && tid < 32 + pb) {
                if (piv >= r0 && piv < r1) rowp((int)(piv - r0))[tid - 32] = krow[tid - 32];
            }
            __syncthreads();
        }
        // ---- elimination of my rows below k, candidates for column k+1: ONE THREAD PER ROW (all 16 panel columns
        //      of a row in one thread: ~60 instructions per row instead of ~100 per quarter row with four lanes per
        //      row -- the panel kernels are bound by the instruction stream of a warp, not by shared-memory bandwidth)
        unsigned long long bkey = 0ull;
        int brow = 0x7fffffff;
        auto eliminate_row = [&](double* row, int li) {
            const double a = row[c];
            const double m = mode == 0 ? a * rpiv : (mode == 1 ? a / pivot : a);
            row[c] = m;
            double nextv = 0.0;
#pragma unroll
            for (int j = 1; j < LU_PB; ++j) {
                if (j > c && j < pb) {                      // (uniform over the block)
                    const double v = fma(-m, pr[j], row[j]);
                    row[j] = v;
                    if (j == c + 1) nextv = v;
                }
            }
            if (has_next) {
                const unsigned long long kk = lp_key(nextv);
                if (kk > bkey) { bkey = kk; brow = (int)(r0 + li); }
            }
        };
        const int lstart = (k + 1 > r0) ? (int)((k + 1 - r0 < nown) ? k + 1 - r0 : nown) : 0;
        for (int li = lstart + tid; li < nres; li += LP_THREADS)             // rows resident in shared memory
            eliminate_row(sm + (size_t)li * LP_PITCH, li);
        // rows beyond the shared-memory capacity (iterates of more than ~1690 rows per SM, e.g. 10^6 rows), in place in
        // global memory: 16 LANES PER ROW, lane = column, so that a row is one coalesced 128-byte access (one thread
        // per row would touch 32 lines per instruction: the 10^6-row LU took 87 ms instead of 25)
        {
            const int col = lane & 15, hw = lane >> 4;
            const bool upd = col > c && col < pb;
            const double prj = upd ? prow[col] : 0.0;
            const int ostart = nres > lstart ? nres : lstart;
            constexpr int UNR = 8;                          // row pairs in flight per warp (the loads are the latency)
            for (int base = ostart + 2 * warp; base < nown; base += 2 * LP_WARPS * UNR) {      // warp-uniform trip count
                double x[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int li = base + u * 2 * LP_WARPS + hw;
                    x[u] = (li < nown && col < pb) ? Ypan[(r0 + li) * ld + col] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int li = base + u * 2 * LP_WARPS + hw;
                    const bool valid = li < nown;
                    const double a = __shfl_sync(0xffffffffu, x[u], (lane & 16) | c);
                    const double m = mode == 0 ? a * rpiv : (mode == 1 ? a / pivot : a);
                    double* row = Ypan + (r0 + (valid ? li : ostart)) * ld;
                    if (valid && col == c) row[col] = m;
                    if (valid && upd) {
                        const double v = fma(-m, prj, x[u]);
                        row[col] = v;
                        if (col == c + 1) {
                            const unsigned long long kk = lp_key(v);
                            if (kk > bkey) { bkey = kk; brow = (int)(r0 + li); }
                        }
                    }
                }
            }
        }
        if (k + 1 < p.pe) publish(k + 1, c + 1, bkey, brow);
    }
    __syncthreads();
    // ---- write my resident rows back
    for (int li = lfirst + rslot; li < nres; li += LP_THREADS / 4) {
        double* dst = Ypan + (r0 + li) * ld;
        const double* src = sm + (size_t)li * LP_PITCH;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = sub + 4 * c;
            if (j < pb) dst[j] = src[j];
        }
    }
    // ---- interchanges of the columns outside the panel, in pivot order (dlaswp): a thread
    //      stays in its column, so the 16 swaps need no synchronisation.  Done by the last CTA.
    if (b == G - 1) {
        for (int j = tid; j < p.l; j += LP_THREADS) {
            if (j >= p.ps && j < p.pe) continue;
            for (int c = 0; c < pb; ++c) {
                const int64_t k = p.ps + c, pv = s_piv[c];
                if (pv != k) {
                    const double t = p.Y[k * ld + j];
                    p.Y[k * ld + j] = p.Y[pv * ld + j];
                    p.Y[pv * ld + j] = t;
                }
            }
        }
    }
    // CL = 1: nobody's shared memory may be released while a peer can still push a record into it --
    // the last push precedes the last cluster barrier, after which only local state is touched.
}

// U12 = L11^{-1} A12 for the panel rows [ps, pe): one thread per trailing column.  The rows are
// updated in place and copied to the small TALL buffer U (pb x (l - pe)) for the GEMM update.
// L11 is staged in shared memory first: read from Y inside the substitution, every load would have to
// wait for the preceding store to the same array (12 us per call for 136 multiply-adds).
__global__ void lu_u12_kernel(double* __restrict__ Y, int64_t ld, int l, int ps, int pe,
                              double* __restrict__ U, int64_t ldu) {
    __shared__ double s_l11[LU_PB][LU_PB + 1];
    const int pb = pe - ps;
    for (int i = threadIdx.x; i < LU_PB * LU_PB; i += blockDim.x) {
        const int r = i / LU_PB, c = i % LU_PB;
        s_l11[r][c] = (r < pb && c < r) ? Y[(int64_t)(ps + r) * ld + ps + c] : 0.0;
    }
    __syncthreads();
    const int j = pe + blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= l) return;
    double u[LU_PB];
#pragma unroll
    for (int r = 0; r < LU_PB; ++r) u[r] = (r < pb) ? Y[(int64_t)(ps + r) * ld + j] : 0.0;
#pragma unroll
    for (int r = 0; r < LU_PB; ++r) {
        if (r < pb) {
            double v = u[r];
#pragma unroll
            for (int c = 0; c < LU_PB; ++c)
                if (c < r) v -= s_l11[r][c] * u[c];
            u[r] = v;
        }
    }
#pragma unroll
    for (int r = 0; r < LU_PB; ++r) {
        if (r < pb) {
            Y[(int64_t)(ps + r) * ld + j] = u[r];
            U[(int64_t)r * ldu + (j - pe)] = u[r];
        }
    }
}

// rows < min(n, l): zero the U part, unit diagonal
__global__ void lu_finalize_kernel(double* __restrict__ Y, int64_t ld, int64_t n, int l) {
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    for (int j = (int)i + threadIdx.x; j < l; j += blockDim.x) Y[i * ld + j] = (j == i) ? 1.0 : 0.0;
}

void lu_reset_flag(gsi_ctx* ctx) {
    GSI_CUDA(cudaMemsetAsync(ctx->dflags, 0, sizeof(int), ctx->stream));          // [0] first zero pivot
}

// Synchronises the stream and throws SingularException if an LU since the last reset met an
// exactly zero pivot (Julia's lu(...; check = true)).
void lu_check_singular(gsi_ctx* ctx) {
    int flag = 0;
    GSI_CUDA(cudaMemcpyAsync(&flag, ctx->dflags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    if (flag != 0) {
        lu_reset_flag(ctx);
        throw Error(GSI_ERR_SINGULAR, "SingularException(" + std::to_string(flag) + "): exactly zero pivot in lu");
    }
}

// All n rows of Y are on this device (single GPU, or the gathered iterate on every rank).
// Enqueues the factorisation; the zero-pivot flag is left for lu_check_singular.
void lu_L_inplace(gsi_ctx* ctx, gsi_buf* Y) {
    GSI_REQUIRE(Y->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "lu: TALL buffer required");
    const int l = (int)Y->cols;
    const int64_t n = Y->rows;
    if (n == 0) return;
    const int kmax = (int)(n < l ? n : l);        // number of elimination steps (n < l: L is n x n, the rest of F.L's columns do not exist)
    int grid = ctx->num_sms * 4;
    const int64_t need = (n + LU_WARPS * 8 - 1) / (LU_WARPS * 8);
    if (grid > need) grid = (int)(need > 0 ? need : 1);
    const size_t xlen = 2 + 2 * (size_t)l;
    const size_t smem = 2 * (size_t)l * sizeof(double);

    // panel driver set-up.  Short iterates (lu.panel = 1): the launch is ONE thread-block cluster of up to
    // 16 CTAs; otherwise (or lu.panel = 2) a cooperative grid of co-resident CTAs, at most one per SM.
    LuPanelParams pp;
    int pgrid = 0, pmode = 0;                 // pmode 1: cluster transport
    size_t psmem = 0;
    auto panel_setup = [&](int mode) {
        pgrid = 0; pmode = mode; psmem = 0;
        const void* kfn = mode ? (const void*)lu_panel_kernel<1> : (const void*)lu_panel_kernel<0>;
        cudaFuncAttributes fa;
        GSI_CUDA(cudaFuncGetAttributes(&fa, kfn));
        const int cap_max = (int)((232448 - fa.sharedSizeBytes) / (LP_PITCH * sizeof(double)));   // 227 KB per CTA, minus the static part
        const int64_t rmin = 256;                                       // below this a CTA's share is not worth a barrier participant
        const int ctas_max = mode ? kPxchClusterMax : ctx->num_sms;
        int64_t R = round_up((n + ctas_max - 1) / ctas_max, 8);
        if (R < rmin) R = rmin;
        if (mode && R > cap_max + 256) return;                          // too long for one cluster: cooperative grid
        const int g = (int)((n + R - 1) / R);
        int cap = (int)R;
        if (cap > cap_max) cap = cap_max;
        psmem = (size_t)cap * LP_PITCH * sizeof(double);
        GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        if (mode) {
            if (g > 8) GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        } else {
            int occ = 0;
            GSI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lu_panel_kernel<0>, LP_THREADS, psmem));
            if (occ < 1 || g > ctx->num_sms * occ || g > kPxchMaxCtas) return;   // per-column driver instead
        }
        pgrid = g;
        pp.Y = Y->d; pp.ld = Y->ld; pp.n = n; pp.l = l;
        pp.recs = ctx->pxch;
        pp.bar = ctx->pbar; pp.bar_base = 0;
        pp.flags = ctx->dflags;
        pp.R = R; pp.cap = cap;
    };
    if (ctx->lu_panel == 1) panel_setup(1);
    if (ctx->lu_panel && pgrid == 0) panel_setup(0);
    if (pgrid > 0 && pmode == 0) GSI_CUDA(cudaMemsetAsync(ctx->pbar, 0, kPbarBytes, ctx->stream));
    auto panel_launch = [&]() -> cudaError_t {
        void* args[] = {&pp};
        if (pmode == 0)
            return cudaLaunchCooperativeKernel((void*)lu_panel_kernel<0>, dim3((unsigned)pgrid), dim3(LP_THREADS), args,
                                               psmem, ctx->stream);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)pgrid);
        cfg.blockDim = dim3(LP_THREADS);
        cfg.dynamicSmemBytes = psmem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)pgrid;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelExC(&cfg, (const void*)lu_panel_kernel<1>, args);
    };
    // per-column driver scratch: cand[grid] | xch[xlen]
    Cand* cand = reinterpret_cast<Cand*>(ctx->scratch);
    double* xch = ctx->scratch + 2 * (size_t)grid;
    GSI_REQUIRE(2 * (size_t)grid + xlen + 16 <= ctx->scratch_doubles, GSI_ERR_UNSUPPORTED, "lu: scratch too small");

    for (int ps = 0; ps < kmax; ps += LU_PB) {
        const int pe = (ps + LU_PB < kmax) ? ps + LU_PB : kmax;
        bool done = false;
        if (pgrid > 0) {
            pp.ps = ps; pp.pe = pe;
            cudaError_t e = panel_launch();
            if (e != cudaSuccess && pmode == 1) {
                // the cluster could not be scheduled (partitioned device, no GPC with that many free SMs):
                // cooperative grid from here on
                cudaGetLastError();
                panel_setup(0);
                if (pgrid > 0) {
                    GSI_CUDA(cudaMemsetAsync(ctx->pbar, 0, kPbarBytes, ctx->stream));
                    pp.ps = ps; pp.pe = pe;
                    e = panel_launch();
                }
            }
            if (pgrid > 0 && e == cudaSuccess) {
                count_launch(ctx);
                pp.bar_base += (unsigned int)(pe - ps);     // one barrier per column step
                done = true;
            } else {
                cudaGetLastError();            // the grid could not be made co-resident (MPS / partitioned device)
                pgrid = 0;
            }
        }
        if (!done) {
            lu_search_kernel<<<grid, LU_THREADS, 0, ctx->stream>>>(Y->d, Y->ld, n, ps, cand);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
            for (int k = ps; k < pe; ++k) {
                lu_pack_kernel<<<1, LU_THREADS, 0, ctx->stream>>>(Y->d, Y->ld, n, l, k, cand, grid, xch);
                lu_eliminate_kernel<<<grid, LU_THREADS, smem, ctx->stream>>>(Y->d, Y->ld, n, l, k, pe, xch, cand,
                                                                              ctx->dflags);
                GSI_CUDA(cudaGetLastError());
                count_launch(ctx, 2);
            }
        }
        if (pe < l) {
            // U12, then the rank-pb update of the trailing matrix
            BufPtr U = make_buf(ctx, GSI_LAYOUT_TALL, pe - ps, l - pe);
            const int ncol = l - pe;
            lu_u12_kernel<<<(ncol + 127) / 128, 128, 0, ctx->stream>>>(Y->d, Y->ld, l, ps, pe, U->d, U->ld);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
            if (pe < n)
                tall_window_update(ctx, Y->d + (int64_t)pe * Y->ld + ps, Y->ld, n - pe, pe - ps, U.get(),
                                   Y->d + (int64_t)pe * Y->ld + pe, Y->ld, -1.0);
        }
    }
    lu_finalize_kernel<<<kmax, 64, 0, ctx->stream>>>(Y->d, Y->ld, n, l);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

}  // namespace gsi
