// Small (l x l, l <= 256) SVD on the device by one-sided (Hestenes) Jacobi -- the
// `svd(B)` of reference src/RandMatFact.jl:86 after B' has been reduced to its l x l
// triangular factor by TSQR.  One warp owns one column pair; the
// l/2 disjoint pairs of a round-robin round run in parallel; rounds are separate
// launches on the stream (the matrix, <= 512 KB, stays in L2).  Columns of M converge
// to U * diag(sigma); they are normalised and sorted (descending, LAPACK order) at the
// end.  High relative accuracy is the reason for Jacobi over bidiagonalisation.
//
// Two drivers of the same rotation sequence (bit-identical results):
//   * jacobi_round_kernel: one launch per round (any size);
//   * jacobi_fused_kernel: ALL sweeps in one launch of a single thread-block cluster (up to
//     8 CTAs x 32 warps = 256 column pairs, i.e. <= 512 columns), rounds separated by the
//     hardware cluster barrier instead of a kernel boundary, convergence decided on the device.
//     A round is ~2 L2 round trips of work, so the ~3 us fixed cost of a launch dominated the
//     per-round version (1881 launches, 10.7 ms at l = 210).  Option "svd.fused" = 2 (and the
//     fall-back of the block driver below, "svd.fused" = 1).
//     (Measured alternative, round 2, removed: the matrix held in the cluster's distributed shared
//     memory instead of L2 -- bit-identical, but SLOWER: 6.5 vs 4.9 ms at l = 210, 5.7 vs 4.4 ms
//     at the 17 472-point case; remote shared-memory round trips do not beat L2 round trips here.)
#include "common.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gsi {

constexpr int JS_WARPS = 8;
constexpr int JF_MAX_CTAS = 8;       // portable cluster size
constexpr int JF_MAX_SWEEPS = 60;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// positions 0..np-1 (np even); round r pairs position t with np-1-t; player at position i:
// i == 0 -> 0, else ((i - 1 + r) mod (np - 1)) + 1.
__device__ __forceinline__ int rr_player(int i, int r, int np) {
    return i == 0 ? 0 : ((i - 1 + r) % (np - 1)) + 1;
}

// The arithmetic of one pair, with every multiply-add spelled out so that both drivers round
// identically whatever the compiler would otherwise contract.
__device__ __forceinline__ void jacobi_acc(double x, double y, double& a, double& b, double& g) {
    a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
}
// 1/x for 1e-280 < |x| < 1e280 and 1/sqrt(x) for 1 <= x < 1e300: hardware seed (MUFU, ~2^-20) + two Newton
// steps, <= 2 ulp.  The scalar chain of a rotation (two divisions, three square roots with the library
// routines: ~1000 cycles) is the critical path of a Jacobi round -- every warp of a round waits for it --
// and the rotation only needs c^2 + s^2 = 1 to rounding, not correctly rounded quotients.
__device__ __forceinline__ double jf_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ double jf_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(__dmul_rn(-h, y), y, 0.5);
    y = fma(y, e, y);
    e = fma(__dmul_rn(-h, y), y, 0.5);
    y = fma(y, e, y);
    return y;
}
__device__ __forceinline__ bool jacobi_angle(double a, double b, double g, double tol, double& c, double& s) {
    const double thr = tol * sqrt(__dmul_rn(a, b));           // independent of the chain below
    const double ag = fabs(g);
    double zeta;                                              // (b - a) / (2 g)
    if (ag > 1e-280 && ag < 1e280) zeta = __dmul_rn(0.5 * (b - a), jf_rcp(g));
    else zeta = (b - a) / (2.0 * g);
    const double az = fabs(zeta);
    double tt;                                                // tan of the rotation angle: sign(zeta) / (|zeta| + sqrt(1 + zeta^2))
    if (az < 1e150) {
        const double w = fma(az, az, 1.0);
        tt = jf_rcp(az + __dmul_rn(w, jf_rsqrt(w)));
    } else {
        tt = 0.5 / az;
    }
    if (zeta < 0.0) tt = -tt;
    c = jf_rsqrt(fma(tt, tt, 1.0));
    s = __dmul_rn(c, tt);
    return !(ag <= thr || g == 0.0);
}
__device__ __forceinline__ void jacobi_rot(double c, double s, double x, double y, double& nx, double& ny) {
    nx = fma(c, x, -__dmul_rn(s, y));
    ny = fma(s, x, __dmul_rn(c, y));
}

// M: column-major, ncols columns of pitch ld.  The rotation angle of a column pair comes from
// its first rows_dot rows; the rotation is applied to all rows_all >= rows_dot rows (rows below
// rows_dot carry a matrix that accumulates the right singular vectors, e.g. an identity).
__global__ void __launch_bounds__(JS_WARPS * 32)
jacobi_round_kernel(double* __restrict__ M, int64_t ld, int rows_dot, int rows_all, int l, int np, int round,
                    double tol, int* __restrict__ rotated) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * JS_WARPS + warp;
    if (t >= np / 2) return;
    int p = rr_player(t, round, np);
    int q = rr_player(np - 1 - t, round, np);
    if (p >= l || q >= l) return;            // dummy player (odd l)
    if (p > q) { const int tmp = p; p = q; q = tmp; }
    double* mp = M + (size_t)p * ld;
    double* mq = M + (size_t)q * ld;
    double a = 0.0, b = 0.0, g = 0.0;
    for (int i = lane; i < rows_dot; i += 32) jacobi_acc(mp[i], mq[i], a, b, g);
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    double c, s;
    if (!jacobi_angle(a, b, g, tol, c, s)) return;
    if (lane == 0) atomicAdd(rotated, 1);
    for (int i = lane; i < rows_all; i += 32) {
        double nx, ny;
        jacobi_rot(c, s, mp[i], mq[i], nx, ny);
        mp[i] = nx;
        mq[i] = ny;
    }
}

// All sweeps in one launch: grid = ONE cluster; pair slot t = cluster-wide warp index.  Matrix
// data moves through L2 (ld.cg / st.cg) and rounds are separated by cluster.sync() (release /
// acquire at cluster scope), so a column written in round r by one CTA is read in round r+1 by
// another.  rotated[sweep] counts the rotations of a sweep; every thread reads it after the
// sweep's last barrier, so the exit decision is uniform.  The arithmetic and its order per pair
// are those of jacobi_round_kernel.  NR > 0: rows_all <= 32*NR and the two columns stay in
// registers between the dot products and the rotation (then ncols <= 256, i.e. <= 128 pairs and
// at most 16 warps per CTA, which leaves 128 registers per thread).
template <int NR>
__global__ void __launch_bounds__(NR > 0 ? 512 : 1024)
jacobi_fused_kernel(double* __restrict__ M, int64_t ld, int rows_dot, int rows_all, int l, int np, double tol,
                    int* __restrict__ rotated, int* __restrict__ sweeps_out) {
    cg::cluster_group cluster = cg::this_cluster();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = (int)cluster.block_rank() * (int)(blockDim.x >> 5) + warp;
    int done = -1;
    for (int sweep = 0; sweep < JF_MAX_SWEEPS; ++sweep) {
        for (int round = 0; round < np - 1; ++round) {
            int p = 0, q = 0;
            bool live = t < np / 2;
            if (live) {
                p = rr_player(t, round, np);
                q = rr_player(np - 1 - t, round, np);
                live = p < l && q < l;                       // dummy player (odd l)
            }
            if (live) {
                if (p > q) { const int tmp = p; p = q; q = tmp; }
                double* mp = M + (size_t)p * ld;
                double* mq = M + (size_t)q * ld;
                double a = 0.0, b = 0.0, g = 0.0;
                double xr[NR > 0 ? NR : 1], yr[NR > 0 ? NR : 1];
                if (NR > 0) {
#pragma unroll
                    for (int k = 0; k < NR; ++k) {
                        const int i = lane + 32 * k;
                        xr[k] = i < rows_all ? __ldcg(mp + i) : 0.0;
                        yr[k] = i < rows_all ? __ldcg(mq + i) : 0.0;
                    }
#pragma unroll
                    for (int k = 0; k < NR; ++k) {
                        if (lane + 32 * k < rows_dot) jacobi_acc(xr[k], yr[k], a, b, g);
                    }
                } else {
                    for (int i = lane; i < rows_dot; i += 32) jacobi_acc(__ldcg(mp + i), __ldcg(mq + i), a, b, g);
                }
                a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
                double c, s;
                if (jacobi_angle(a, b, g, tol, c, s)) {
                    if (lane == 0) atomicAdd(rotated + sweep, 1);
                    if (NR > 0) {
#pragma unroll
                        for (int k = 0; k < NR; ++k) {
                            const int i = lane + 32 * k;
                            if (i < rows_all) {
                                double nx, ny;
                                jacobi_rot(c, s, xr[k], yr[k], nx, ny);
                                __stcg(mp + i, nx);
                                __stcg(mq + i, ny);
                            }
                        }
                    } else {
                        for (int i = lane; i < rows_all; i += 32) {
                            double nx, ny;
                            jacobi_rot(c, s, __ldcg(mp + i), __ldcg(mq + i), nx, ny);
                            __stcg(mp + i, nx);
                            __stcg(mq + i, ny);
                        }
                    }
                }
            }
            cluster.sync();
        }
        if (__ldcg(rotated + sweep) == 0) { done = sweep + 1; break; }
    }
    if (t == 0 && lane == 0) *sweeps_out = done;
}

// Block one-sided Jacobi, all sweeps in ONE launch of a single cluster of C CTAs (option "svd.fused" = 1,
// the default where it fits).  The columns form 2C blocks of bs columns.  A sweep is
//   (a) CTA c orthogonalises the pairs INSIDE blocks 2c and 2c+1 (round-robin inside each block), then
//   (b) 2C-1 block rounds (round-robin over the blocks): CTA c loads its two blocks (I, J) into SHARED
//       MEMORY, rotates the bs x bs cross pairs (bs inner rounds of bs disjoint pairs, one warp per
//       pair, __syncthreads between inner rounds) and writes the blocks back,
// i.e. every column pair exactly once per sweep (a cyclic ordering, like the flat drivers above) but
// with ONE cluster barrier per bs inner rounds instead of one per round, and the column data served
// from shared memory instead of L2: l = 210 runs ~16 barriers + 16 block loads per sweep instead of 209
// barriers with two L2 round trips each.  The rotation arithmetic is that of the flat drivers; the
// ordering differs, so results agree with them to rounding, not bit for bit.
constexpr int JB_MAX_BS = 32;        // columns per block: one warp per cross pair of an inner round

// round-robin player without an integer division (0 <= i < np, 0 <= r < np - 1)
__device__ __forceinline__ int rr_player_fast(int i, int r, int np) {
    if (i == 0) return 0;
    int x = i - 1 + r;
    if (x >= np - 1) x -= np - 1;
    return x + 1;
}

// One column pair of the shared tile (pitch-rpad columns mp, mq; mp is the lower global index).  NR > 0:
// rpad = 32 * NR, rows >= rows_all are zero, and the two columns stay in registers between the dot
// products and the rotation (straight-line code: a single warp's dependent instruction stream is what
// an inner round costs -- the generic-loop version issued ~440 instructions per pair, 3600 cycles).
template <int NR>
__device__ __forceinline__ int jb_rotate_pair(double* __restrict__ mp, double* __restrict__ mq, int rows_dot, int rows_all,
                                              double tol, int lane) {
    double aa = 0.0, bb = 0.0, gg = 0.0, cs, sn;
    if constexpr (NR > 0) {
        double x[NR], y[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k) { x[k] = mp[lane + 32 * k]; y[k] = mq[lane + 32 * k]; }
#pragma unroll
        for (int k = 0; k < NR; ++k)
            if (lane + 32 * k < rows_dot) jacobi_acc(x[k], y[k], aa, bb, gg);
        aa = warp_sum(aa); bb = warp_sum(bb); gg = warp_sum(gg);
        if (!jacobi_angle(aa, bb, gg, tol, cs, sn)) return 0;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            double nx, ny;
            jacobi_rot(cs, sn, x[k], y[k], nx, ny);
            mp[lane + 32 * k] = nx;
            mq[lane + 32 * k] = ny;
        }
        return 1;
    } else {
        for (int i = lane; i < rows_dot; i += 32) jacobi_acc(mp[i], mq[i], aa, bb, gg);
        aa = warp_sum(aa); bb = warp_sum(bb); gg = warp_sum(gg);
        if (!jacobi_angle(aa, bb, gg, tol, cs, sn)) return 0;
        for (int i = lane; i < rows_all; i += 32) {
            double nx, ny;
            jacobi_rot(cs, sn, mp[i], mq[i], nx, ny);
            mp[i] = nx;
            mq[i] = ny;
        }
        return 1;
    }
}

// blockDim = 32 * max(bs, 2): warp w owns cross pair w of an inner round.
// The column blocks never return to global memory between stages: a CTA works on a shared-memory tile
// (two blocks) and, when a stage is done, PUSHES each of its blocks into the tile of the CTA that owns it in
// the next stage (distributed shared memory, double-buffered tiles) -- one cluster barrier per stage and no
// global round trip (the first version stored / fenced / reloaded the blocks through L2: about half of its
// time at l = 110).  Global memory is read once at the start and written once at the end.
template <int NR>
__global__ void __launch_bounds__(NR >= 4 ? 512 : 1024)
jacobi_block_kernel(double* __restrict__ M, int64_t ld, int rows_dot, int rows_all, int l, int bs, int rpad, double tol,
                    int* __restrict__ rotated, int* __restrict__ sweeps_out) {
    extern __shared__ double xs_all[];               // [2 tiles][2 * bs][rpad]: block I columns, then block J columns
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks(), c = (int)cluster.block_rank();
    const int nblk = 2 * C;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (int)(blockDim.x >> 5);
    const int tile = 2 * bs * rpad;
    const int rcopy = NR > 0 ? rpad : rows_all;      // rows moved per column (the zero pad rows travel along)
    // valid columns of block blk (columns >= l do not exist)
    auto ncols_of = [&](int blk) -> int {
        const int left = l - blk * bs;
        return left < 0 ? 0 : (left < bs ? left : bs);
    };
    // stage -1: intra-block stage (CTA c owns blocks 2c, 2c+1); stage br >= 0: cross stage br of the round-robin.
    // Owner CTA and tile half of block X in a stage:
    auto owner_of = [&](int X, int stage, int& cta, int& half) {
        if (stage < 0) { cta = X >> 1; half = X & 1; return; }
        int i = 0;                                   // position of X in round `stage`
        if (X != 0) {
            int x = (X - 1 - stage) % (nblk - 1);
            if (x < 0) x += nblk - 1;
            i = x + 1;
        }
        const int partner = rr_player(nblk - 1 - i, stage, nblk);
        cta = i < C ? i : nblk - 1 - i;
        half = X < partner ? 0 : 1;
    };
    for (int i = threadIdx.x; i < 2 * tile; i += blockDim.x) xs_all[i] = 0.0;      // pad rows stay zero for good
    __syncthreads();
    int cur = 0;
    // ---- the only read of the matrix: my two blocks of the first (intra) stage
    for (int half = 0; half < 2; ++half) {
        const int blk = 2 * c + half, nc = ncols_of(blk);
        for (int j = warp; j < nc; j += nwarps) {
            const double* src = M + (size_t)(blk * bs + j) * ld;
            double* dst = xs_all + (size_t)(half * bs + j) * rpad;
            for (int i = lane; i < rows_all; i += 32) dst[i] = __ldcg(src + i);
        }
    }
    cluster.sync();                                  // every CTA of the cluster runs (and has zeroed its tiles) before anybody pushes into it
    // push the block in tile half `half` of my current tile to its owner in `next_stage`
    auto push_block = [&](int blk, int half, int next_stage) {
        const int nc = ncols_of(blk);
        if (nc == 0) return;
        int cta, dhalf;
        owner_of(blk, next_stage, cta, dhalf);
        const double* src0 = xs_all + (size_t)cur * tile + (size_t)half * bs * rpad;
        double* dst0 = cluster.map_shared_rank(xs_all + (size_t)(cur ^ 1) * tile + (size_t)dhalf * bs * rpad, cta);
        for (int j = warp; j < nc; j += nwarps)
            for (int i = lane; i < rcopy; i += 32) dst0[(size_t)j * rpad + i] = src0[(size_t)j * rpad + i];
    };
    int done = -1;
    for (int sweep = 0; sweep < JF_MAX_SWEEPS; ++sweep) {
        int nrot = 0;
        // ---- (a) pairs inside my two blocks
        {
            double* xs = xs_all + (size_t)cur * tile;
            const int I = 2 * c, J = 2 * c + 1;
            const int ncI = ncols_of(I), ncJ = ncols_of(J);
            const int npb = (bs + 1) / 2 * 2;                    // players per block (a dummy if bs is odd)
            for (int round = 0; round < npb - 1; ++round) {
                for (int slot = warp; slot < npb; slot += nwarps) {        // npb / 2 pairs per block, two blocks
                    const int half = slot >= npb / 2;
                    const int t = slot - half * (npb / 2);
                    int p = rr_player_fast(t, round, npb), q = rr_player_fast(npb - 1 - t, round, npb);
                    if (p > q) { const int tmp = p; p = q; q = tmp; }
                    if (q < (half ? ncJ : ncI))
                        nrot += jb_rotate_pair<NR>(xs + (size_t)(half * bs + p) * rpad, xs + (size_t)(half * bs + q) * rpad,
                                                   rows_dot, rows_all, tol, lane);
                }
                __syncthreads();
            }
            push_block(I, 0, 0);
            push_block(J, 1, 0);
            cluster.sync();
            cur ^= 1;
        }
        // ---- (b) cross pairs of the block pairs, round-robin over the 2C blocks
        for (int br = 0; br < nblk - 1; ++br) {
            double* xs = xs_all + (size_t)cur * tile;
            int I = rr_player(c, br, nblk), J = rr_player(nblk - 1 - c, br, nblk);
            if (I > J) { const int tmp = I; I = J; J = tmp; }
            const int ncI = ncols_of(I), ncJ = ncols_of(J);
            int jq = warp;                                       // inner round r pairs column w of I with column (w + r) mod bs of J
            for (int r = 0; r < bs; ++r) {
                if (warp < ncI && jq < ncJ)
                    nrot += jb_rotate_pair<NR>(xs + (size_t)warp * rpad, xs + (size_t)(bs + jq) * rpad, rows_dot, rows_all, tol,
                                               lane);
                if (++jq >= bs) jq -= bs;
                __syncthreads();
            }
            const int next_stage = (br + 1 < nblk - 1) ? br + 1 : -1;
            push_block(I, 0, next_stage);
            push_block(J, 1, next_stage);
            if (br + 1 == nblk - 1 && lane == 0 && nrot) atomicAdd(rotated + sweep, nrot);   // rotations of this sweep
            cluster.sync();
            cur ^= 1;
        }
        // ---- convergence: rotations of this sweep over the whole cluster (the adds precede the last barrier)
        if (__ldcg(rotated + sweep) == 0) { done = sweep + 1; break; }
    }
    // ---- the only write: after the last stage every CTA holds blocks 2c, 2c+1 again
    {
        const double* xs = xs_all + (size_t)cur * tile;
        for (int half = 0; half < 2; ++half) {
            const int blk = 2 * c + half, nc = ncols_of(blk);
            for (int j = warp; j < nc; j += nwarps) {
                double* dst = M + (size_t)(blk * bs + j) * ld;
                const double* src = xs + (size_t)(half * bs + j) * rpad;
                for (int i = lane; i < rows_all; i += 32) dst[i] = src[i];
            }
        }
    }
    if (c == 0 && threadIdx.x == 0) *sweeps_out = done;
}

// sigma_j = ||m_j||, rank by descending sigma (ties: lower index first), write normalised
// columns to U in sorted order.
__global__ void jacobi_finalize_kernel(const double* __restrict__ M, int l, double* __restrict__ U,
                                       double* __restrict__ sigma) {
    __shared__ double s_sig[kMaxWideCols];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int j = warp; j < l; j += nw) {
        double a = 0.0;
        for (int i = lane; i < l; i += 32) { const double x = M[(size_t)j * l + i]; a += x * x; }
        a = warp_sum(a);
        if (lane == 0) s_sig[j] = sqrt(a);
    }
    __syncthreads();
    for (int j = warp; j < l; j += nw) {
        const double sj = s_sig[j];
        int rank = 0;
        for (int i = 0; i < l; ++i) {
            const double si = s_sig[i];
            if (si > sj || (si == sj && i < j)) ++rank;
        }
        const double inv = sj > 0.0 ? 1.0 / sj : 0.0;
        for (int i = lane; i < l; i += 32) U[(size_t)rank * l + i] = M[(size_t)j * l + i] * inv;
        if (lane == 0) sigma[rank] = sj;
    }
}

// One-sided Jacobi sweeps (round-robin ordering, dgesvj tolerance sqrt(rows)*eps) until a whole
// sweep rotates nothing: afterwards the first rows_dot rows of M hold U * diag(sigma) column by
// column and rows [rows_dot, rows_all) hold the input rows there multiplied by the accumulated
// rotations V.
template <int NR>
static cudaError_t launch_jacobi_fused(gsi_ctx* ctx, int nctas, int wpc, double* M, int64_t ld, int rows_dot, int rows_all,
                                int ncols, int np, double tol) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nctas);
    cfg.blockDim = dim3((unsigned)wpc * 32);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)nctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, jacobi_fused_kernel<NR>, M, ld, rows_dot, rows_all, ncols, np, tol, ctx->jflags,
                              ctx->jflags + JF_MAX_SWEEPS);
}

void svd_check(gsi_ctx* ctx);

// Block driver (svd.fused = 1); returns false when the two column blocks of a CTA do not fit its shared memory
// or the cluster cannot be launched.
static bool jacobi_sweeps_block(gsi_ctx* ctx, double* M, int64_t ld, int rows_dot, int rows_all, int ncols, double tol,
                                bool defer_check) {
    if (ncols < 4) return false;
    int C = JF_MAX_CTAS;
    while (C > 1 && 2 * C > ncols) C >>= 1;                  // at least one column per block
    const int bs = (ncols + 2 * C - 1) / (2 * C);
    if (bs > JB_MAX_BS) return false;
    const int wpc = bs < 2 ? 2 : bs;                         // one warp per cross pair of an inner round
    // rows per lane held in registers: 2 / 4 / 8 (then the column pitch is 32 * NR, zero padded); 0: streamed
    int NR = rows_all <= 64 ? 2 : (rows_all <= 128 && wpc <= 16) ? 4 : (rows_all <= 256 && wpc <= 16) ? 8 : 0;
    const int rpad = NR > 0 ? 32 * NR : (rows_all + 3) / 4 * 4 + 4;      // column pitch in shared memory
    const size_t smem = (size_t)2 * 2 * bs * rpad * sizeof(double);     // two tiles (double-buffered exchange)
    if (smem > (size_t)200 * 1024) return false;
    const void* kfn = NR == 2 ? (const void*)jacobi_block_kernel<2> : NR == 4 ? (const void*)jacobi_block_kernel<4>
                    : NR == 8 ? (const void*)jacobi_block_kernel<8> : (const void*)jacobi_block_kernel<0>;
    GSI_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GSI_CUDA(cudaMemsetAsync(ctx->jflags, 0, (JF_MAX_SWEEPS + 1) * sizeof(int), ctx->stream));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)C);
    cfg.blockDim = dim3((unsigned)wpc * 32);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int* rot = ctx->jflags;
    int* swp = ctx->jflags + JF_MAX_SWEEPS;
    void* args[] = {&M, &ld, &rows_dot, &rows_all, &ncols, (void*)&bs, (void*)&rpad, &tol, &rot, &swp};
    const cudaError_t e = cudaLaunchKernelExC(&cfg, kfn, args);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    count_launch(ctx);
    ctx->svd_pending = true;
    if (!defer_check) svd_check(ctx);
    return true;
}

// Single-launch driver; returns false when the problem does not fit one cluster.
static bool jacobi_sweeps_fused(gsi_ctx* ctx, double* M, int64_t ld, int rows_dot, int rows_all, int ncols, int np,
                                double tol, bool defer_check) {
    const int pairs = np / 2;
    if (pairs > JF_MAX_CTAS * 32 || ncols < 2) return false;
    int nctas = pairs < JF_MAX_CTAS ? pairs : JF_MAX_CTAS;
    const int wpc = (pairs + nctas - 1) / nctas;
    GSI_CUDA(cudaMemsetAsync(ctx->jflags, 0, (JF_MAX_SWEEPS + 1) * sizeof(int), ctx->stream));
    cudaError_t e;
    if (rows_all <= 128 && wpc <= 16) e = launch_jacobi_fused<4>(ctx, nctas, wpc, M, ld, rows_dot, rows_all, ncols, np, tol);
    else if (rows_all <= 256 && wpc <= 16) e = launch_jacobi_fused<8>(ctx, nctas, wpc, M, ld, rows_dot, rows_all, ncols, np, tol);
    else e = launch_jacobi_fused<0>(ctx, nctas, wpc, M, ld, rows_dot, rows_all, ncols, np, tol);
    if (e != cudaSuccess) {
        // the cluster could not be scheduled on this device / partition: use the per-round driver from now on
        cudaGetLastError();
        ctx->svd_fused = 0;
        return false;
    }
    count_launch(ctx);
    ctx->svd_pending = true;
    if (!defer_check) svd_check(ctx);
    return true;
}

// Convergence verdict of the last single-launch Jacobi run whose check was deferred (the randsvd
// driver defers it to its one synchronisation point).  Synchronises the stream.
void svd_check(gsi_ctx* ctx) {
    if (!ctx->svd_pending) return;
    ctx->svd_pending = false;
    int h = 0;
    GSI_CUDA(cudaMemcpyAsync(&h, ctx->jflags + JF_MAX_SWEEPS, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->svd_last_sweeps = h;
    GSI_REQUIRE(h > 0, GSI_ERR_NO_CONVERGENCE, "Jacobi SVD did not converge in 60 sweeps");
}

void jacobi_sweeps(gsi_ctx* ctx, double* M, int64_t ld, int rows_dot, int rows_all, int ncols, bool defer_check) {
    const int np = (ncols + 1) / 2 * 2;
    if (ctx->svd_fused == 1 &&
        jacobi_sweeps_block(ctx, M, ld, rows_dot, rows_all, ncols, sqrt((double)rows_dot) * 2.220446049250313e-16, defer_check))
        return;
    if (ctx->svd_fused &&
        jacobi_sweeps_fused(ctx, M, ld, rows_dot, rows_all, ncols, np, sqrt((double)rows_dot) * 2.220446049250313e-16,
                            defer_check))
        return;
    const int nrounds = np - 1;
    const int blocks = (np / 2 + JS_WARPS - 1) / JS_WARPS;
    int* rotated = ctx->dflags + 1;
    const double tol = sqrt((double)rows_dot) * 2.220446049250313e-16;   // dgesvj's sqrt(m)*eps
    const int max_sweeps = 60;
    bool converged = (ncols == 1);
    for (int sweep = 0; sweep < max_sweeps && !converged; ++sweep) {
        GSI_CUDA(cudaMemsetAsync(rotated, 0, sizeof(int), ctx->stream));
        for (int r = 0; r < nrounds; ++r) {
            jacobi_round_kernel<<<blocks, JS_WARPS * 32, 0, ctx->stream>>>(M, ld, rows_dot, rows_all, ncols, np, r, tol,
                                                                           rotated);
        }
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, nrounds);
        int h = 0;
        GSI_CUDA(cudaMemcpyAsync(&h, rotated, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
        converged = (h == 0);
    }
    GSI_REQUIRE(converged, GSI_ERR_NO_CONVERGENCE, "Jacobi SVD did not converge in 60 sweeps");
}

// M: l x l column-major (ld = l), device.  U (l x l, ld = l) and sigma (l) device outputs.
void svd_small(gsi_ctx* ctx, double* M, int l, double* U, double* sigma, bool defer_check) {
    GSI_REQUIRE(l >= 1 && l <= kMaxWideCols, GSI_ERR_UNSUPPORTED, "svd_small: l must be in 1..1024");
    jacobi_sweeps(ctx, M, l, l, l, l, defer_check);
    jacobi_finalize_kernel<<<1, 1024, 0, ctx->stream>>>(M, l, U, sigma);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

}  // namespace gsi
