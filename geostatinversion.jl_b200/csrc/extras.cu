// Adaptive range finder (reference src/RandMatFact.jl:15-48, Halko et al. Alg 4.2) and
// eig_nystrom (reference src/RandMatFact.jl:92-102, Alg 5.5).
//
// The adaptive finder is BLAS-2 by construction (one operator pass per basis vector:
// 8*n^2 bytes per step for a dense A -> HBM-bound), so it is built from small strided
// vector kernels around op_apply with a single column; the stopping test needs one scalar
// on the host per step.  Yfull / Qfull are kept column-major on the device because the
// basis may outgrow the 256-column TALL limit while it is being built; `omegas` and `Q_out` may
// be TALL (<= 256 columns) or COLMAJOR (any width, e.g. the default maxvec = min(m, n)).
#include "common.cuh"
#include "algos.h"

#define GSI_API extern "C" __attribute__((visibility("default")))

namespace gsi {

__device__ __forceinline__ double ex_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ double ex_block_sum(double v) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = ex_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
    r = ex_warp_sum(r);
    return r;
}

// c[k] = sum_i M[i, k] * y[i*ys]   for k in [0, ncols); M column-major (ld)
__global__ void colmat_t_vec_kernel(const double* __restrict__ M, int64_t ld, int64_t n, const double* __restrict__ y,
                                    int64_t ys, double* __restrict__ c) {
    const int k = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += M[(int64_t)k * ld + i] * y[i * ys];
    s = ex_block_sum(s);
    if (threadIdx.x == 0) c[k] = s;
}

// out[i*os] = y[i*ys] - sum_k M[i,k] c[k]
__global__ void vec_minus_colmat_vec_kernel(const double* __restrict__ M, int64_t ld, int64_t n, int ncols,
                                            const double* __restrict__ c, const double* __restrict__ y, int64_t ys,
                                            double* __restrict__ out, int64_t os) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = 0.0;
    for (int k = 0; k < ncols; ++k) s += M[(int64_t)k * ld + i] * c[k];
    out[i * os] = y[i * ys] - s;
}

// norms[k] = ||M[:, k]||_2 for k in [0, ncols)
__global__ void col_norms_kernel(const double* __restrict__ M, int64_t ld, int64_t n, double* __restrict__ norms) {
    const int k = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double v = M[(int64_t)k * ld + i]; s += v * v; }
    s = ex_block_sum(s);
    if (threadIdx.x == 0) norms[k] = sqrt(s);
}

// q = y / ||y||   (norm supplied in nrm[0])
__global__ void scale_by_inv_norm_kernel(const double* __restrict__ y, int64_t n, const double* __restrict__ nrm,
                                         double* __restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) q[i] += (1.0 / nrm[0]) * y[i];          // axpy!(1/norm(Yj), Yj, Qj) onto zeros (:34)
}

// for each of ncols columns Yi of Y: Yi -= dot(q, Yi) * q       (:42-45)
__global__ void reorth_cols_kernel(double* __restrict__ Y, int64_t ld, int64_t n, const double* __restrict__ q) {
    double* yi = Y + (int64_t)blockIdx.x * ld;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += q[i] * yi[i];
    s = ex_block_sum(s);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) yi[i] += -s * q[i];
}

__global__ void strided_copy_kernel(const double* __restrict__ src, int64_t ss, double* __restrict__ dst, int64_t ds,
                                    int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i * ds] = src[i * ss];
}

static inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

void rangefinder_adaptive(gsi_op* op, const gsi_buf* Omega0, const gsi_buf* omegas, double epsilon, int64_t r,
                          gsi_buf* Q_out, int64_t* j_out) {
    gsi_ctx* ctx = op->ctx;
    GSI_REQUIRE(ctx->world == 1, GSI_ERR_UNSUPPORTED, "adaptive rangefinder is single-GPU");
    const int64_t m = op->m, n = op->n;
    GSI_REQUIRE(m == n, GSI_ERR_UNSUPPORTED,
                "adaptive rangefinder needs a square operator (the reference allocates Yfull with n rows, RandMatFact.jl:18)");
    GSI_REQUIRE(Omega0->layout == GSI_LAYOUT_TALL && Omega0->rows == n && Omega0->cols == r, GSI_ERR_DIMENSION_MISMATCH,
                "Omega0 must be TALL n x r");
    GSI_REQUIRE(omegas->rows == n, GSI_ERR_DIMENSION_MISMATCH, "omegas must have n rows");
    const int64_t maxvec = omegas->cols;
    GSI_REQUIRE(Q_out->rows == m && Q_out->cols == maxvec, GSI_ERR_DIMENSION_MISMATCH, "Q_out must be m x maxvec");
    GSI_REQUIRE((size_t)(maxvec + r + 2) <= ctx->scratch_doubles, GSI_ERR_UNSUPPORTED,
                "adaptive rangefinder: maxvec + r exceeds the scratch area");
    // element (i, c) of a TALL (row pitch ld) or COLMAJOR (column pitch ld) buffer: base + i * rs + c * cs
    auto strides = [](const gsi_buf* b, int64_t& rs, int64_t& cs) {
        if (b->layout == GSI_LAYOUT_TALL) { rs = b->ld; cs = 1; } else { rs = 1; cs = b->ld; }
    };
    int64_t om_rs, om_cs, qo_rs, qo_cs;
    strides(omegas, om_rs, om_cs);
    strides(Q_out, qo_rs, qo_cs);
    cudaStream_t st = ctx->stream;
    const int64_t ycols = r + maxvec;
    // column-major Yfull (n x (r+maxvec)) and Qfull (m x maxvec), zero initialised (:18, :23)
    BufPtr Yf = make_buf(ctx, GSI_LAYOUT_COLMAJOR, n, ycols);
    BufPtr Qf = make_buf(ctx, GSI_LAYOUT_COLMAJOR, m, maxvec);
    BufPtr colin = make_buf(ctx, GSI_LAYOUT_TALL, n, 1);
    BufPtr colout = make_buf(ctx, GSI_LAYOUT_TALL, m, 1);
    BufPtr Y0 = make_buf(ctx, GSI_LAYOUT_TALL, m, r);
    double* small = ctx->scratch;                 // c[maxvec] | norms[r] | nrm[1] | yj[n] kept separately
    double* cvec = small;
    double* norms = small + maxvec;
    double* nrm = norms + r + 1;
    BufPtr yj = make_buf(ctx, GSI_LAYOUT_COLMAJOR, n, 1);
    // Y = A * randn(n, r)                                                         (:19-20)
    op_apply(op, 0, Omega0, Y0.get());
    for (int64_t c = 0; c < r; ++c)
        strided_copy_kernel<<<nblk(n), 256, 0, st>>>(Y0->d + c, Y0->ld, Yf->d + c * Yf->ld, 1, n);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx, (int)r);
    const double thresh = epsilon / sqrt(200.0 / M_PI);
    std::vector<double> hn(r);
    int64_t j = 0;
    while (true) {
        col_norms_kernel<<<(unsigned)r, 256, 0, st>>>(Yf->d + j * Yf->ld, Yf->ld, n, norms);       // :26
        GSI_CUDA(cudaMemcpyAsync(hn.data(), norms, r * sizeof(double), cudaMemcpyDeviceToHost, st));
        GSI_CUDA(cudaStreamSynchronize(st));
        count_launch(ctx);
        double mx = 0.0;
        for (double v : hn) mx = v > mx ? v : mx;
        if (!(mx > thresh)) break;
        if (j >= maxvec) {
            *j_out = j;
            throw Error(GSI_ERR_NO_CONVERGENCE, "adaptive rangefinder: maxvec basis vectors did not reach epsilon");
        }
        j += 1;                                                                              // :27
        const double* Ycol = Yf->d + (j - 1) * Yf->ld;
        double* Qj = Qf->d + (j - 1) * Qf->ld;
        // QtYj = Q' Yj ; Yj = Yj - Q QtYj (local copy) ; Qj = Yj / ||Yj||                 (:28-34)
        if (j > 1) colmat_t_vec_kernel<<<(unsigned)(j - 1), 256, 0, st>>>(Qf->d, Qf->ld, m, Ycol, 1, cvec);
        vec_minus_colmat_vec_kernel<<<nblk(m), 256, 0, st>>>(Qf->d, Qf->ld, m, (int)(j - 1), cvec, Ycol, 1, yj->d, 1);
        col_norms_kernel<<<1, 256, 0, st>>>(yj->d, yj->ld, m, nrm);
        scale_by_inv_norm_kernel<<<nblk(m), 256, 0, st>>>(yj->d, m, nrm, Qj);
        // Aomega = A * omega_j                                                              (:36-37)
        strided_copy_kernel<<<nblk(n), 256, 0, st>>>(omegas->d + (j - 1) * om_cs, om_rs, colin->d, colin->ld, n);
        op_apply(op, 0, colin.get(), colout.get());
        // ynew = Aomega - Q (Q' Aomega) ; Yfull[:, r+j] = ynew                              (:38-40)
        colmat_t_vec_kernel<<<(unsigned)j, 256, 0, st>>>(Qf->d, Qf->ld, m, colout->d, colout->ld, cvec);
        vec_minus_colmat_vec_kernel<<<nblk(m), 256, 0, st>>>(Qf->d, Qf->ld, m, (int)j, cvec, colout->d, colout->ld,
                                                             Yf->d + (r + j - 1) * Yf->ld, 1);
        // for i = j+1 : j+r-1: Yi -= dot(Qj, Yi) Qj                                          (:42-45)
        if (r > 1) reorth_cols_kernel<<<(unsigned)(r - 1), 256, 0, st>>>(Yf->d + j * Yf->ld, Yf->ld, n, Qj);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, 8);
    }
    // Qfull[:, 1:j] -> output (columns beyond j are zero)
    GSI_CUDA(cudaMemsetAsync(Q_out->d, 0, Q_out->bytes(), st));
    for (int64_t c = 0; c < j; ++c)
        strided_copy_kernel<<<nblk(m), 256, 0, st>>>(Qf->d + c * Qf->ld, 1, Q_out->d + c * qo_cs, qo_rs, m);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx, (int)j);
    GSI_CUDA(cudaStreamSynchronize(st));
    *j_out = j;
}

// ---- blocked adaptive range finder: opt-in, NOT the parity mode -------------
// out[c] = ||Y[:, c]||_2 of a TALL buffer (one CTA per column)
__global__ void tall_col_norms_kernel(const double* __restrict__ Y, int64_t ld, int64_t n, double* __restrict__ norms) {
    const int c = blockIdx.x;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) { const double v = Y[i * ld + c]; s += v * v; }
    s = ex_block_sum(s);
    if (threadIdx.x == 0) norms[c] = sqrt(s);
}
// Y -= T on the first `cols` columns of two TALL buffers of equal pitch
__global__ void tall_sub_kernel(double* __restrict__ Y, const double* __restrict__ T, int64_t ld, int64_t n, int cols) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t i = idx / cols;
    const int c = (int)(idx - i * cols);
    if (i < n) Y[i * ld + c] -= T[i * ld + c];
}

// HMT algorithm 4.2 with the random vectors consumed a BLOCK at a time: Y = A * Omega_b is one
// tensor-core GEMM pass over A instead of b GEMV passes (8 n^2 bytes per vector -> per block),
// the projection against the basis found so far is block Gram-Schmidt with re-orthogonalisation
// (two pairs of skinny GEMMs), the block is orthonormalised by the Householder QR (then projected
// and orthonormalised once more, for blocks that overshoot the rank), and the
// stopping test is the reference's estimator evaluated on the b fresh probes of a block:
//     max_c ||(I - Q Q') A omega_c|| <= epsilon / (10 sqrt(2/pi)).
// It draws the vectors in a different order of use than src/RandMatFact.jl:15-48 (all b columns
// of a block join the basis together), so the basis -- and its size, rounded up to the block --
// differ from the reference's: ||A - Q Q' A|| meets the same bound, the result is not bit-parity.
void rangefinder_adaptive_blocked(gsi_op* op, const gsi_buf* omegas, double epsilon, int64_t block, gsi_buf* Q_out,
                                  int64_t* j_out) {
    gsi_ctx* ctx = op->ctx;
    GSI_REQUIRE(ctx->world == 1, GSI_ERR_UNSUPPORTED, "adaptive rangefinder is single-GPU");
    const int64_t m = op->m, n = op->n;
    GSI_REQUIRE(block >= 1 && block <= kMaxCols && block <= m, GSI_ERR_INVALID_ARGUMENT,
                "blocked adaptive rangefinder: 1 <= block <= min(256, m) required");
    GSI_REQUIRE(omegas->rows == n, GSI_ERR_DIMENSION_MISMATCH, "omegas must have n rows");
    const int64_t maxvec = omegas->cols;
    GSI_REQUIRE(Q_out->rows == m && Q_out->cols == maxvec, GSI_ERR_DIMENSION_MISMATCH, "Q_out must be m x maxvec");
    cudaStream_t st = ctx->stream;
    int64_t om_rs, om_cs, qo_rs, qo_cs;
    if (omegas->layout == GSI_LAYOUT_TALL) { om_rs = omegas->ld; om_cs = 1; } else { om_rs = 1; om_cs = omegas->ld; }
    if (Q_out->layout == GSI_LAYOUT_TALL) { qo_rs = Q_out->ld; qo_cs = 1; } else { qo_rs = 1; qo_cs = Q_out->ld; }
    BufPtr Qf = make_buf(ctx, GSI_LAYOUT_COLMAJOR, m, maxvec);
    const double thresh = epsilon / sqrt(200.0 / M_PI);
    double* norms = ctx->scratch;
    std::vector<double> hn((size_t)block);
    int64_t j = 0;
    while (true) {
        const int64_t b = (maxvec - j < block) ? maxvec - j : block;
        if (b <= 0) {
            *j_out = j;
            throw Error(GSI_ERR_NO_CONVERGENCE, "adaptive rangefinder: maxvec basis vectors did not reach epsilon");
        }
        BufPtr Om = make_buf(ctx, GSI_LAYOUT_TALL, n, b);
        for (int64_t c = 0; c < b; ++c)
            strided_copy_kernel<<<nblk(n), 256, 0, st>>>(omegas->d + (j + c) * om_cs, om_rs, Om->d + c, Om->ld, n);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, (int)b);
        BufPtr Y = make_buf(ctx, GSI_LAYOUT_TALL, m, b);
        op_apply(op, 0, Om.get(), Y.get());                                     // Y = A * Omega_b
        // Y -= Q (Q' Y), `passes` times (block Gram-Schmidt against the basis found so far)
        auto project = [&](int passes) {
            if (j == 0) return;
            gsi_buf qv = *Qf;
            qv.owns = false; qv.cols = j;
            BufPtr C = make_buf(ctx, GSI_LAYOUT_TALL, j, b);
            BufPtr T = make_buf(ctx, GSI_LAYOUT_TALL, m, b);
            for (int pass = 0; pass < passes; ++pass) {
                dense_apply(ctx, &qv, 1, Y.get(), C.get(), 1.0);
                dense_apply(ctx, &qv, 0, C.get(), T.get(), 1.0);
                tall_sub_kernel<<<nblk(m * b), 256, 0, st>>>(Y->d, T->d, Y->ld, m, (int)b);
                GSI_CUDA(cudaGetLastError());
                count_launch(ctx, 3);
            }
        };
        project(2);
        tall_col_norms_kernel<<<(unsigned)b, 256, 0, st>>>(Y->d, Y->ld, m, norms);
        GSI_CUDA(cudaMemcpyAsync(hn.data(), norms, b * sizeof(double), cudaMemcpyDeviceToHost, st));
        GSI_CUDA(cudaStreamSynchronize(st));
        count_launch(ctx);
        double mx = 0.0;
        for (int64_t c = 0; c < b; ++c) mx = hn[c] > mx ? hn[c] : mx;
        if (!(mx > thresh)) break;
        lu_reset_flag(ctx);
        qr_thinQ_inplace(ctx, Y.get(), nullptr);                                // orthonormalise the block
        // a block that overshoots the rank has columns of pure rounding noise whose directions are not
        // orthogonal to the basis: project the orthonormalised block once more and re-orthonormalise
        if (j > 0) {
            project(1);
            qr_thinQ_inplace(ctx, Y.get(), nullptr);
        }
        for (int64_t c = 0; c < b; ++c)
            strided_copy_kernel<<<nblk(m), 256, 0, st>>>(Y->d + c, Y->ld, Qf->d + (j + c) * Qf->ld, 1, m);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, (int)b);
        j += b;
    }
    GSI_CUDA(cudaMemsetAsync(Q_out->d, 0, Q_out->bytes(), st));
    for (int64_t c = 0; c < j; ++c)
        strided_copy_kernel<<<nblk(m), 256, 0, st>>>(Qf->d + c * Qf->ld, 1, Q_out->d + c * qo_cs, qo_rs, m);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx, (int)j);
    GSI_CUDA(cudaStreamSynchronize(st));
    *j_out = j;
}

// ---- small dense kernels for eig_nystrom (l <= 256, single CTA, matrix column-major ld = l)
// upper Cholesky B = C' C of the upper triangle of B (dpotrf 'U'); flag = first failing pivot
__global__ void chol_upper_kernel(double* __restrict__ B, int l, int* __restrict__ flag) {
    __shared__ double s_piv;
    for (int k = 0; k < l; ++k) {
        if (threadIdx.x == 0) {
            const double d = B[(size_t)k * l + k];
            if (!(d > 0.0)) { if (*flag == 0) *flag = k + 1; s_piv = 0.0; }
            else { s_piv = sqrt(d); B[(size_t)k * l + k] = s_piv; }
        }
        __syncthreads();
        const double piv = s_piv;
        if (piv == 0.0) return;
        for (int j = k + 1 + threadIdx.x; j < l; j += blockDim.x) B[(size_t)j * l + k] /= piv;   // row k of C
        __syncthreads();
        // trailing update: B[i, j] -= C[k, i] * C[k, j]  for k < i <= j
        for (int idx = threadIdx.x; idx < (l - k - 1) * (l - k - 1); idx += blockDim.x) {
            const int i = k + 1 + idx % (l - k - 1), j = k + 1 + idx / (l - k - 1);
            if (i <= j) B[(size_t)j * l + i] -= B[(size_t)i * l + k] * B[(size_t)j * l + k];
        }
        __syncthreads();
    }
    // zero the strict lower triangle
    for (int idx = threadIdx.x; idx < l * l; idx += blockDim.x) {
        const int i = idx % l, j = idx / l;
        if (i > j) B[idx] = 0.0;
    }
}

// Cinv = inv(C) for upper-triangular C (dtrtri), column j by back substitution
__global__ void triu_inverse_kernel(const double* __restrict__ C, int l, double* __restrict__ Cinv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= l) return;
    // solve C x = e_j ; x has non-zeros in rows 0..j
    for (int i = l - 1; i > j; --i) Cinv[(size_t)j * l + i] = 0.0;
    for (int i = j; i >= 0; --i) {
        double s = (i == j) ? 1.0 : 0.0;
        for (int k = i + 1; k <= j; ++k) s -= C[(size_t)k * l + i] * Cinv[(size_t)j * l + k];
        Cinv[(size_t)j * l + i] = s / C[(size_t)i * l + i];
    }
}

__global__ void tall_head_to_cm_kernel(const double* __restrict__ T, int64_t ld, int l, double* __restrict__ M) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l * l) return;
    const int r = idx % l, c = idx / l;
    M[idx] = T[(int64_t)r * ld + c];
}

void eig_nystrom(gsi_op* op, const gsi_buf* Q, gsi_buf* U_out, double* Sigma_host) {
    gsi_ctx* ctx = op->ctx;
    GSI_REQUIRE(ctx->world == 1, GSI_ERR_UNSUPPORTED, "eig_nystrom is single-GPU");
    GSI_REQUIRE(Q->layout == GSI_LAYOUT_TALL && Q->rows == op->n, GSI_ERR_DIMENSION_MISMATCH, "Q must be TALL with size(A,2) rows");
    GSI_REQUIRE(op->m == op->n, GSI_ERR_DIMENSION_MISMATCH, "eig_nystrom needs a square operator");
    const int l = (int)Q->cols;
    GSI_REQUIRE(U_out->layout == GSI_LAYOUT_TALL && U_out->rows == op->m && U_out->cols == l, GSI_ERR_DIMENSION_MISMATCH,
                "U_out must be TALL n x l");
    cudaStream_t st = ctx->stream;
    BufPtr B1 = make_buf(ctx, GSI_LAYOUT_TALL, op->m, l);
    op_apply(op, 0, Q, B1.get());                                            // B1 = A * Q       (:93)
    // B2 = Q' * B1 (l x l): Q as the column-major matrix Q' (ld x n) times B1             (:94)
    gsi_buf qview;
    qview.ctx = ctx; qview.layout = GSI_LAYOUT_COLMAJOR; qview.rows = l; qview.cols = Q->rows; qview.ld = Q->ld;
    qview.d = Q->d; qview.owns = false;
    BufPtr B2t = make_buf(ctx, GSI_LAYOUT_TALL, l, l);
    dense_apply(ctx, &qview, 0, B1.get(), B2t.get(), 1.0);
    double* small = static_cast<double*>(pool_alloc(ctx, ((size_t)4 * l * l + l) * 8));
    struct Guard { gsi_ctx* c; void* p; size_t b; ~Guard() { pool_free(c, p, b); } } guard{ctx, small, ((size_t)4 * l * l + l) * 8};
    double* C = small;
    double* Cinv = small + (size_t)l * l;
    double* R = small + (size_t)2 * l * l;
    double* Us = small + (size_t)3 * l * l;
    double* sig = small + (size_t)4 * l * l;
    tall_head_to_cm_kernel<<<(l * l + 255) / 256, 256, 0, st>>>(B2t->d, B2t->ld, l, C);
    GSI_CUDA(cudaMemsetAsync(ctx->dflags + 2, 0, sizeof(int), st));
    chol_upper_kernel<<<1, 1024, 0, st>>>(C, l, ctx->dflags + 2);            // cholesky(Hermitian(B2)).U   (:95)
    int flag = 0;
    GSI_CUDA(cudaMemcpyAsync(&flag, ctx->dflags + 2, sizeof(int), cudaMemcpyDeviceToHost, st));
    GSI_CUDA(cudaStreamSynchronize(st));
    count_launch(ctx, 2);
    if (flag != 0) throw Error(GSI_ERR_NOT_POSDEF, "PosDefException(" + std::to_string(flag) + "): Q'AQ is not positive definite");
    triu_inverse_kernel<<<(l + 63) / 64, 64, 0, st>>>(C, l, Cinv);           // inv(C)                       (:96)
    BufPtr Ct = make_buf(ctx, GSI_LAYOUT_TALL, l, l);
    small_cm_to_tall(ctx, Cinv, l, l, l, Ct.get());
    BufPtr F = make_buf(ctx, GSI_LAYOUT_TALL, op->m, l);
    tall_times_small(ctx, B1.get(), Ct.get(), F.get());                      // F = B1 * inv(C)
    // U, Sigma = svd(F): F = Q_F R, R = U_R S V'  =>  U = Q_F U_R                          (:97)
    tsqr_thinQ(op, F.get(), false, R);
    svd_small(ctx, R, l, Us, sig);
    BufPtr Ut = make_buf(ctx, GSI_LAYOUT_TALL, l, l);
    small_cm_to_tall(ctx, Us, l, l, l, Ut.get());
    tall_times_small(ctx, F.get(), Ut.get(), U_out);
    GSI_CUDA(cudaMemcpyAsync(Sigma_host, sig, (size_t)l * 8, cudaMemcpyDeviceToHost, st));
    GSI_CUDA(cudaStreamSynchronize(st));
    count_launch(ctx, 2);
}

}  // namespace gsi

using namespace gsi;

template <typename F>
static int32_t guarded(F&& f) {
    try { f(); return GSI_OK; }
    catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error(e.what()); return GSI_ERR_INVALID_ARGUMENT; }
    catch (...) { set_last_error("unknown error"); return GSI_ERR_INVALID_ARGUMENT; }
}

GSI_API int32_t gsi_rangefinder_adaptive(gsi_op* op, const gsi_buf* Omega0, const gsi_buf* omegas, double epsilon,
                                         int64_t r, gsi_buf* Q_out, int64_t* j_out) {
    return guarded([&] {
        GSI_REQUIRE(op && Omega0 && omegas && Q_out && j_out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_REQUIRE(r >= 1, GSI_ERR_INVALID_ARGUMENT, "adaptive rangefinder: r >= 1 required");
        GSI_CUDA(cudaSetDevice(op->ctx->device));
        *j_out = 0;
        rangefinder_adaptive(op, Omega0, omegas, epsilon, r, Q_out, j_out);
    });
}

GSI_API int32_t gsi_rangefinder_adaptive_blocked(gsi_op* op, const gsi_buf* omegas, double epsilon, int64_t block,
                                                 gsi_buf* Q_out, int64_t* j_out) {
    return guarded([&] {
        GSI_REQUIRE(op && omegas && Q_out && j_out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(op->ctx->device));
        *j_out = 0;
        rangefinder_adaptive_blocked(op, omegas, epsilon, block, Q_out, j_out);
    });
}

GSI_API int32_t gsi_eig_nystrom(gsi_op* op, const gsi_buf* Q, gsi_buf* U_out, double* Sigma_host) {
    return guarded([&] {
        GSI_REQUIRE(op && Q && U_out && Sigma_host, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(op->ctx->device));
        eig_nystrom(op, Q, U_out, Sigma_host);
    });
}
