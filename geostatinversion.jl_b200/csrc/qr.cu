// Tall-skinny Householder QR with explicit thin Q:
// replaces `Matrix(qr(Y, Val(true)).Q)` (reference src/RandMatFact.jl:57-58,75-76 ->
// LAPACK dgeqp3 + dorgqr; only range(Q) matters) and the QR of B' that
// precedes the small SVD (`svd(B)`, :86).
//
// Local factorisation: LAPACK dgeqr2/dlarfg reflectors, one column per step, each step =
//   qr_house  (1 CTA)  reduce the dot products  g_j = sum_{i>k} Y[i,k] Y[i,j]  that the
//                      previous update accumulated, form (beta, tau, 1/(alpha-beta)),
//                      update row k, publish tau*w_j
//   qr_update (grid)   one read+write pass over the trailing rows: scale column k to v,
//                      apply the reflector, and -- fused -- accumulate the dot products
//                      the NEXT column needs (warp-shuffle broadcast of the column-(k+1)
//                      entry, per-lane partial sums, one block reduction at the end).
// Q is then formed in place by the same two-kernel pattern run backwards (dorg2r).
// Across GPUs the R factors are all-gathered and re-factored redundantly (TSQR).
#include "common.cuh"

namespace gsi {

constexpr int QR_THREADS = 1024;     // 32 warps/SM: the update pass is latency-bound with fewer
constexpr int QR_WARPS = QR_THREADS / 32;
constexpr int QR_MAXC = (kMaxCols + 31) / 32;   // column chunks of 32 per lane

// partial[b][j] (b = CTA) -> reduced in the *_house kernels
struct QrScal { double tau, scale, beta, pad; };

__device__ __forceinline__ double block_reduce_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < QR_WARPS) r = sh[threadIdx.x];
    if (warp == 0) {
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        if (lane == 0) sh[0] = r;
    }
    __syncthreads();
    r = sh[0];
    __syncthreads();
    return r;
}

// Shared accumulation epilogue: psum[c] holds this lane's partial for column j0 + lane + 32c.
__device__ __forceinline__ void store_partials(const double (&psum)[QR_MAXC], int j0, int l, double* sm /*[QR_WARPS][l]*/,
                                               double* __restrict__ partial_row) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < QR_MAXC; ++c) {
        const int j = j0 + lane + 32 * c;
        if (j < l) sm[warp * l + j] = psum[c];
    }
    __syncthreads();
    for (int j = j0 + threadIdx.x; j < l; j += QR_THREADS) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < QR_WARPS; ++w) s += sm[w * l + j];
        partial_row[j] = s;
    }
}

// g_j = sum_{i > 0} Y[i,0] * Y[i,j]  (initial dot products for column 0)
__global__ void __launch_bounds__(QR_THREADS)
qr_dots0_kernel(const double* __restrict__ Y, int64_t ld, int64_t n, int l, double* __restrict__ partial) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double psum[QR_MAXC];
#pragma unroll
    for (int c = 0; c < QR_MAXC; ++c) psum[c] = 0.0;
    for (int64_t i = 1 + (int64_t)blockIdx.x * QR_WARPS + warp; i < n; i += (int64_t)gridDim.x * QR_WARPS) {
        const double* yrow = Y + i * ld;
        const double y0 = yrow[0];
#pragma unroll
        for (int c = 0; c < QR_MAXC; ++c) {
            const int j = lane + 32 * c;
            if (j < l) psum[c] += y0 * yrow[j];
        }
    }
    store_partials(psum, 0, l, sm, partial + (size_t)blockIdx.x * l);
}

constexpr int QH_THREADS = 256;      // single-CTA scalar kernels
// Householder scalars of column k + row-k update.  tw[j] = tau * w_j for j > k.
__global__ void __launch_bounds__(QH_THREADS)
qr_house_kernel(double* __restrict__ Y, int64_t ld, int l, int k, const double* __restrict__ partial, int nparts,
                double* __restrict__ tw, double* __restrict__ taus, QrScal* __restrict__ scal) {
    __shared__ double s_g[kMaxCols];
    __shared__ double s_tau, s_scale;
    for (int j = k + threadIdx.x; j < l; j += QH_THREADS) {
        double s = 0.0;
        for (int b = 0; b < nparts; ++b) s += partial[(size_t)b * l + j];
        s_g[j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double alpha = Y[(int64_t)k * ld + k];
        const double xnorm2 = s_g[k];
        double tau = 0.0, scale = 0.0, beta = alpha;
        if (xnorm2 > 0.0) {                         // dlarfg
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
    for (int j = k + 1 + threadIdx.x; j < l; j += QH_THREADS) {
        const double ykj = Y[(int64_t)k * ld + j];
        const double w = ykj + scale * s_g[j];      // v' * Y[:, j]   (v_k = 1)
        const double t = tau * w;
        Y[(int64_t)k * ld + j] = ykj - t;
        tw[j] = t;
    }
}

// rows i > k: Y[i,k] <- v_i = scale*Y[i,k];  Y[i,j] -= v_i*tw[j];  accumulate next dots.
__global__ void __launch_bounds__(QR_THREADS)
qr_update_kernel(double* __restrict__ Y, int64_t ld, int64_t n, int l, int k, const double* __restrict__ tw,
                 const QrScal* __restrict__ scal, double* __restrict__ partial) {
    extern __shared__ double sm[];      // [QR_WARPS][l] + tw copy [l]
    double* s_tw = sm + QR_WARPS * l;
    for (int j = threadIdx.x; j < l; j += QR_THREADS) s_tw[j] = (j > k) ? tw[j] : 0.0;
    __syncthreads();
    const double scale = scal->scale;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double psum[QR_MAXC];
#pragma unroll
    for (int c = 0; c < QR_MAXC; ++c) psum[c] = 0.0;
    const int nchunks = (l - k + 31) / 32;
    for (int64_t i = k + 1 + (int64_t)blockIdx.x * QR_WARPS + warp; i < n; i += (int64_t)gridDim.x * QR_WARPS) {
        double* yrow = Y + i * ld;
        const double v = scale * yrow[k];
        double nv[QR_MAXC];
#pragma unroll
        for (int c = 0; c < QR_MAXC; ++c) {
            nv[c] = 0.0;
            if (c < nchunks) {
                const int j = k + lane + 32 * c;
                if (j < l) {
                    nv[c] = (j == k) ? v : yrow[j] - v * s_tw[j];
                    yrow[j] = nv[c];
                }
            }
        }
        // column k+1 entry lives in lane 1 of chunk 0
        const double ynext = __shfl_sync(0xffffffffu, nv[0], 1);
        if (i > k + 1) {
#pragma unroll
            for (int c = 0; c < QR_MAXC; ++c) psum[c] += ynext * nv[c];
        }
    }
    store_partials(psum, k, l, sm, partial + (size_t)blockIdx.x * l);
}

// ---- explicit Q (dorg2r), backwards ---------------------------------------------------------
// Step k: d_j = sum_{i>k} v_k[i] Q[i,j] arrives in `partial` (rows i > k; accumulated by the
// previous org_update, i.e. of step k+1).
__global__ void __launch_bounds__(QH_THREADS)
org_house_kernel(double* __restrict__ Y, int64_t ld, int l, int k, const double* __restrict__ partial, int nparts,
                 const double* __restrict__ taus, double* __restrict__ tw) {
    const double tau = taus[k];
    for (int j = k + 1 + threadIdx.x; j < l; j += QH_THREADS) {
        double d = 0.0;
        for (int b = 0; b < nparts; ++b) d += partial[(size_t)b * l + j];
        const double qkj = Y[(int64_t)k * ld + j];
        const double w = qkj + d;
        const double t = tau * w;
        Y[(int64_t)k * ld + j] = qkj - t;
        tw[j] = t;
    }
    // column k above the diagonal holds R entries: Q has zeros there
    for (int i = threadIdx.x; i < k; i += QH_THREADS) Y[(int64_t)i * ld + k] = 0.0;
    if (threadIdx.x == 0) Y[(int64_t)k * ld + k] = 1.0 - tau;
}

// rows i > k: Q[i,j] -= v_i*tw[j] (j > k), Q[i,k] = -tau*v_i; accumulate dots with v_{k-1}.
__global__ void __launch_bounds__(QR_THREADS)
org_update_kernel(double* __restrict__ Y, int64_t ld, int64_t n, int l, int k, const double* __restrict__ tw,
                  const double* __restrict__ taus, double* __restrict__ partial) {
    extern __shared__ double sm[];
    double* s_tw = sm + QR_WARPS * l;
    for (int j = threadIdx.x; j < l; j += QR_THREADS) s_tw[j] = (j > k) ? tw[j] : 0.0;
    __syncthreads();
    const double tau = taus[k];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double psum[QR_MAXC];
#pragma unroll
    for (int c = 0; c < QR_MAXC; ++c) psum[c] = 0.0;
    const int nchunks = (l - k + 31) / 32;
    for (int64_t i = k + 1 + (int64_t)blockIdx.x * QR_WARPS + warp; i < n; i += (int64_t)gridDim.x * QR_WARPS) {
        double* yrow = Y + i * ld;
        const double v = yrow[k];
        const double vprev = (k > 0) ? yrow[k - 1] : 0.0;
#pragma unroll
        for (int c = 0; c < QR_MAXC; ++c) {
            if (c < nchunks) {
                const int j = k + lane + 32 * c;
                if (j < l) {
                    const double nv = (j == k) ? -tau * v : yrow[j] - v * s_tw[j];
                    yrow[j] = nv;
                    psum[c] += vprev * nv;
                }
            }
        }
    }
    store_partials(psum, k, l, sm, partial + (size_t)blockIdx.x * l);
}

// add row k's own term  v_{k-1}[k] * Q[k, j]  (j >= k) to partial slot 0 for step k-1
__global__ void org_rowterm_kernel(const double* __restrict__ Y, int64_t ld, int l, int k, double* __restrict__ partial) {
    const double vk = Y[(int64_t)k * ld + (k - 1)];
    for (int j = k + threadIdx.x; j < l; j += blockDim.x) partial[j] += vk * Y[(int64_t)k * ld + j];
}

__global__ void extract_R_kernel(const double* __restrict__ Y, int64_t ld, int l, double* __restrict__ R) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= l * l) return;
    const int r = idx % l, c = idx / l;
    R[(size_t)c * l + r] = (r <= c) ? Y[(int64_t)r * ld + c] : 0.0;
}

void qr_thinQ_inplace(gsi_ctx* ctx, gsi_buf* Y, double* Rdev) {
    GSI_REQUIRE(Y->layout == GSI_LAYOUT_TALL, GSI_ERR_INVALID_ARGUMENT, "qr: TALL buffer required");
    const int l = (int)Y->cols;
    const int64_t n = Y->rows;
    GSI_REQUIRE(n >= l, GSI_ERR_UNSUPPORTED, "qr: fewer (local) rows than columns is not supported");
    int grid = ctx->num_sms;
    const int64_t need = (n + QR_WARPS - 1) / QR_WARPS;
    if (grid > need) grid = (int)(need > 0 ? need : 1);
    // scratch: partial[grid*l] | tw[l] | taus[l] | scal
    const size_t need_doubles = (size_t)grid * l + 2 * (size_t)l + 8;
    GSI_REQUIRE(need_doubles <= ctx->scratch_doubles, GSI_ERR_UNSUPPORTED, "qr: scratch too small");
    double* partial = ctx->scratch;
    double* tw = partial + (size_t)grid * l;
    double* taus = tw + l;
    QrScal* scal = reinterpret_cast<QrScal*>(taus + l);
    const size_t smem_upd = ((size_t)QR_WARPS * l + l) * sizeof(double);
    const size_t smem_dot = (size_t)QR_WARPS * l * sizeof(double);
    cudaStream_t st = ctx->stream;
    GSI_CUDA(cudaFuncSetAttribute(qr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    GSI_CUDA(cudaFuncSetAttribute(org_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    GSI_CUDA(cudaFuncSetAttribute(qr_dots0_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dot));

    qr_dots0_kernel<<<grid, QR_THREADS, smem_dot, st>>>(Y->d, Y->ld, n, l, partial);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
    for (int k = 0; k < l; ++k) {
        qr_house_kernel<<<1, QH_THREADS, 0, st>>>(Y->d, Y->ld, l, k, partial, grid, tw, taus, scal);
        qr_update_kernel<<<grid, QR_THREADS, smem_upd, st>>>(Y->d, Y->ld, n, l, k, tw, scal, partial);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, 2);
    }
    if (Rdev) {
        extract_R_kernel<<<(l * l + 255) / 256, 256, 0, st>>>(Y->d, Y->ld, l, Rdev);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
    // ---- form Q in place, k = l-1 .. 0
    GSI_CUDA(cudaMemsetAsync(partial, 0, (size_t)grid * l * sizeof(double), st));
    for (int k = l - 1; k >= 0; --k) {
        org_house_kernel<<<1, QH_THREADS, 0, st>>>(Y->d, Y->ld, l, k, partial, grid, taus, tw);
        org_update_kernel<<<grid, QR_THREADS, smem_upd, st>>>(Y->d, Y->ld, n, l, k, tw, taus, partial);
        if (k > 0) org_rowterm_kernel<<<1, QH_THREADS, 0, st>>>(Y->d, Y->ld, l, k, partial);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, k > 0 ? 3 : 2);
    }
}

}  // namespace gsi
