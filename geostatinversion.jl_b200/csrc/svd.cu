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
//     per-round version (1881 launches, 10.7 ms at l = 210).  Option "svd.fused".
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
__device__ __forceinline__ bool jacobi_angle(double a, double b, double g, double tol, double& c, double& s) {
    if (fabs(g) <= tol * sqrt(__dmul_rn(a, b)) || g == 0.0) return false;
    const double zeta = (b - a) / (2.0 * g);
    const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(fma(zeta, zeta, 1.0)));
    c = 1.0 / sqrt(fma(tt, tt, 1.0));
    s = __dmul_rn(c, tt);
    return true;
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
    GSI_REQUIRE(h > 0, GSI_ERR_NO_CONVERGENCE, "Jacobi SVD did not converge in 60 sweeps");
}

void jacobi_sweeps(gsi_ctx* ctx, double* M, int64_t ld, int rows_dot, int rows_all, int ncols, bool defer_check) {
    const int np = (ncols + 1) / 2 * 2;
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
