// Small (l x l, l <= 256) SVD on the device by one-sided (Hestenes) Jacobi -- the
// `svd(B)` of reference src/RandMatFact.jl:86 after B' has been reduced to its l x l
// triangular factor by TSQR.  One warp owns one column pair; the
// l/2 disjoint pairs of a round-robin round run in parallel; rounds are separate
// launches on the stream (the matrix, <= 512 KB, stays in L2).  Columns of M converge
// to U * diag(sigma); they are normalised and sorted (descending, LAPACK order) at the
// end.  High relative accuracy is the reason for Jacobi over bidiagonalisation.
#include "common.cuh"

namespace gsi {

constexpr int JS_WARPS = 8;

__device__ __forceinline__ double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// positions 0..np-1 (np even); round r pairs position t with np-1-t; player at position i:
// i == 0 -> 0, else ((i - 1 + r) mod (np - 1)) + 1.
__device__ __forceinline__ int rr_player(int i, int r, int np) {
    return i == 0 ? 0 : ((i - 1 + r) % (np - 1)) + 1;
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
    for (int i = lane; i < rows_dot; i += 32) {
        const double x = mp[i], y = mq[i];
        a += x * x; b += y * y; g += x * y;
    }
    a = warp_sum(a); b = warp_sum(b); g = warp_sum(g);
    if (fabs(g) <= tol * sqrt(a * b) || g == 0.0) return;
    if (lane == 0) atomicAdd(rotated, 1);
    const double zeta = (b - a) / (2.0 * g);
    const double tt = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double c = 1.0 / sqrt(1.0 + tt * tt);
    const double s = c * tt;
    for (int i = lane; i < rows_all; i += 32) {
        const double x = mp[i], y = mq[i];
        mp[i] = c * x - s * y;
        mq[i] = s * x + c * y;
    }
}

// sigma_j = ||m_j||, rank by descending sigma (ties: lower index first), write normalised
// columns to U in sorted order.
__global__ void jacobi_finalize_kernel(const double* __restrict__ M, int l, double* __restrict__ U,
                                       double* __restrict__ sigma) {
    __shared__ double s_sig[kMaxCols];
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
void jacobi_sweeps(gsi_ctx* ctx, double* M, int64_t ld, int rows_dot, int rows_all, int ncols) {
    const int np = (ncols + 1) / 2 * 2;
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
void svd_small(gsi_ctx* ctx, double* M, int l, double* U, double* sigma) {
    GSI_REQUIRE(l >= 1 && l <= kMaxCols, GSI_ERR_UNSUPPORTED, "svd_small: l must be in 1..256");
    jacobi_sweeps(ctx, M, l, l, l, l);
    jacobi_finalize_kernel<<<1, 1024, 0, ctx->stream>>>(M, l, U, sigma);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

}  // namespace gsi
