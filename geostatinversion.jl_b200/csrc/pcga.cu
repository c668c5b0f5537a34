// PCGA / RGA device helpers:
//   * PCGALowRankMatrix mat-vec (reference src/lowrank.jl:83-97)
//   * IterativeSolvers.lsqr on that operator (reference call site src/lsqr.jl:54), the whole
//     Paige-Saunders loop in ONE single-CTA kernel (the system is (nobs+1)^2 <= ~500^2:
//     latency-bound, so one launch, vectors in shared memory, warp-shuffle dot reductions)
//   * the parameter update s = X*beta + sum_i xi_i * dot(eta_i, xi_bar) (src/lsqr.jl:55-61)
//   * the `paramstorun` batch (src/lsqr.jl:37-43)
//   * rga's sketch products S*V and S*R*S' (src/GeostatInversion.jl:102) on the DMMA GEMM.
//   * pcgadirect's dense solve (reference src/direct.jl:49-58): HQH = E E' on the DMMA GEMM,
//     x = pinv([HQH + R, HX; HX', 0]) b by one-sided Jacobi with accumulated right vectors and
//     Julia's pinv cut-off.
#include "common.cuh"
#include "algos.h"

#define GSI_API extern "C" __attribute__((visibility("default")))

namespace gsi {

constexpr int LS_THREADS = 1024;
constexpr int LS_WARPS = LS_THREADS / 32;

struct PcgaOp {
    const double* E;      // nobs x K column-major (ld = lde)
    const double* HX;     // nobs
    const double* Rdiag;  // nobs or null
    const double* Rdense; // nobs x nobs column-major or null
    int nobs, K;
    int64_t lde, ldr;
};

__device__ __forceinline__ double ls_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; red[] has LS_WARPS entries.  All threads get the result.
__device__ double ls_block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = ls_warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (lane < LS_WARPS) ? red[lane] : 0.0;
    r = ls_warp_sum(r);
    return r;
}

// out = A x for the saddle-point operator; x, out have nobs+1 entries (shared or global).
// d (K entries) is scratch.  All threads of the CTA must call.
__device__ void pcga_matvec(const PcgaOp& A, const double* x, double* out, double* d, double* red) {
    const int nobs = A.nobs, K = A.K;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    // d_i = dot(eta_i, xshort)  (one warp per eta);  slot K: dot(HX, xshort)
    for (int i = warp; i <= K; i += LS_WARPS) {
        const double* col = (i < K) ? A.E + (int64_t)i * A.lde : A.HX;
        double s = 0.0;
        for (int j = lane; j < nobs; j += 32) s += col[j] * x[j];
        s = ls_warp_sum(s);
        if (lane == 0) d[i] = s;
    }
    __syncthreads();
    const double xend = x[nobs];
    for (int j = threadIdx.x; j < nobs; j += LS_THREADS) {
        double v;
        if (A.Rdiag) v = A.Rdiag[j] * x[j];
        else {
            v = 0.0;
            for (int c = 0; c < nobs; ++c) v += A.Rdense[(int64_t)c * A.ldr + j] * x[c];
        }
        for (int i = 0; i < K; ++i) v += A.E[(int64_t)i * A.lde + j] * d[i];   // += eta_i[j] * dotp
        v += A.HX[j] * xend;
        out[j] = v;
    }
    if (threadIdx.x == 0) out[nobs] = d[K];
    __syncthreads();
}

__global__ void __launch_bounds__(LS_THREADS) pcga_matvec_kernel(PcgaOp A, const double* x, double* v) {
    extern __shared__ double sm[];
    double* d = sm;
    double* red = sm + A.K + 1;
    pcga_matvec(A, x, v, d, red);
}

struct LsqrOut { double itn, istop, Anorm, Acond, rnorm, Arnorm, xnorm, pad; };

// IterativeSolvers.lsqr(A, b) with x0 = 0, damp = 0 (restated from Paige & Saunders in the
// IterativeSolvers 0.9 form, see oracle/lsqr.py for the statement-by-statement version).
__global__ void __launch_bounds__(LS_THREADS)
pcga_lsqr_kernel(PcgaOp A, const double* __restrict__ b, double atol, double btol, double conlim, int maxiter,
                 double* __restrict__ xout, LsqrOut* __restrict__ info) {
    extern __shared__ double sm[];
    const int m = A.nobs + 1;
    double* u = sm;
    double* v = u + m;
    double* w = v + m;
    double* x = w + m;
    double* tmp = x + m;
    double* d = tmp + m;
    double* red = d + A.K + 1;
    const int tid = threadIdx.x;

    for (int j = tid; j < m; j += LS_THREADS) { x[j] = 0.0; v[j] = 0.0; u[j] = b[j]; }   // u = b - A*0
    __syncthreads();
    double part = 0.0;
    for (int j = tid; j < m; j += LS_THREADS) part += u[j] * u[j];
    double beta = sqrt(ls_block_sum(part, red));
    double alpha = 0.0;
    if (beta > 0.0) {
        const double ib = 1.0 / beta;
        for (int j = tid; j < m; j += LS_THREADS) u[j] *= ib;
        pcga_matvec(A, u, v, d, red);                       // v = A' u  (A symmetric)
        part = 0.0;
        for (int j = tid; j < m; j += LS_THREADS) part += v[j] * v[j];
        alpha = sqrt(ls_block_sum(part, red));
    }
    if (alpha > 0.0) {
        const double ia = 1.0 / alpha;
        for (int j = tid; j < m; j += LS_THREADS) v[j] *= ia;
    }
    __syncthreads();
    for (int j = tid; j < m; j += LS_THREADS) w[j] = v[j];
    __syncthreads();

    int itn = 0, istop = 0;
    double Anorm = 0.0, Acond = 0.0, ddnorm = 0.0, res2 = 0.0, xnorm = 0.0, xxnorm = 0.0, z = 0.0, sn2 = 0.0;
    double cs2 = -1.0;
    const double ctol = conlim > 0.0 ? 1.0 / conlim : 0.0;
    double Arnorm = alpha * beta;
    double rhobar = alpha, phibar = beta, rnorm = beta;
    const double bnorm = beta;
    if (Arnorm != 0.0) {
        while (itn < maxiter && istop == 0) {
            ++itn;
            pcga_matvec(A, v, tmp, d, red);                 // tmp = A v
            for (int j = tid; j < m; j += LS_THREADS) u[j] = -alpha * u[j] + tmp[j];
            __syncthreads();
            part = 0.0;
            for (int j = tid; j < m; j += LS_THREADS) part += u[j] * u[j];
            beta = sqrt(ls_block_sum(part, red));
            if (beta > 0.0) {
                const double ib = 1.0 / beta;
                for (int j = tid; j < m; j += LS_THREADS) u[j] *= ib;
                Anorm = sqrt(Anorm * Anorm + alpha * alpha + beta * beta);
                pcga_matvec(A, u, tmp, d, red);             // tmp = A' u
                for (int j = tid; j < m; j += LS_THREADS) v[j] = -beta * v[j] + tmp[j];
                __syncthreads();
                part = 0.0;
                for (int j = tid; j < m; j += LS_THREADS) part += v[j] * v[j];
                alpha = sqrt(ls_block_sum(part, red));
                if (alpha > 0.0) {
                    const double ia = 1.0 / alpha;
                    for (int j = tid; j < m; j += LS_THREADS) v[j] *= ia;
                }
                __syncthreads();
            }
            // plane rotations (damp = 0)
            const double rhobar1 = sqrt(rhobar * rhobar);
            const double cs1 = rhobar / rhobar1;
            const double psi = 0.0;
            phibar = cs1 * phibar;
            const double rho = sqrt(rhobar1 * rhobar1 + beta * beta);
            const double cs = rhobar1 / rho;
            const double sn = beta / rho;
            const double theta = sn * alpha;
            rhobar = -cs * alpha;
            const double phi = cs * phibar;
            phibar = sn * phibar;
            const double tau = sn * phi;
            const double t1 = phi / rho;
            const double t2 = -theta / rho;
            const double irho = 1.0 / rho;
            part = 0.0;
            for (int j = tid; j < m; j += LS_THREADS) {
                x[j] += t1 * w[j];
                const double wn = t2 * w[j] + v[j];
                w[j] = wn;
                const double wr = wn * irho;
                part += wr * wr;
            }
            ddnorm += sqrt(ls_block_sum(part, red));        // IterativeSolvers: ddnorm += norm(w/rho)
            const double delta = sn2 * rho;
            const double gambar = -cs2 * rho;
            const double rhs = phi - delta * z;
            const double zbar = rhs / gambar;
            xnorm = sqrt(xxnorm + zbar * zbar);
            const double gamma = sqrt(gambar * gambar + theta * theta);
            cs2 = gambar / gamma;
            sn2 = theta / gamma;
            z = rhs / gamma;
            xxnorm += z * z;
            Acond = Anorm * sqrt(ddnorm);
            const double res1 = phibar * phibar;
            res2 = res2 + psi * psi;
            rnorm = sqrt(res1 + res2);
            Arnorm = alpha * fabs(tau);
            const double test1 = rnorm / bnorm;
            const double test2 = (Anorm * rnorm != 0.0) ? Arnorm / (Anorm * rnorm) : INFINITY;
            const double test3 = (Acond != 0.0) ? 1.0 / Acond : INFINITY;
            const double tt1 = test1 / (1.0 + Anorm * xnorm / bnorm);
            const double rtol = btol + atol * Anorm * xnorm / bnorm;
            if (itn >= maxiter) istop = 7;
            if (1.0 + test3 <= 1.0) istop = 6;
            if (1.0 + test2 <= 1.0) istop = 5;
            if (1.0 + tt1 <= 1.0) istop = 4;
            if (test3 <= ctol) istop = 3;
            if (test2 <= atol) istop = 2;
            if (test1 <= rtol) istop = 1;
        }
    }
    __syncthreads();
    for (int j = tid; j < m; j += LS_THREADS) xout[j] = x[j];
    if (tid == 0) {
        info->itn = itn; info->istop = istop; info->Anorm = Anorm; info->Acond = Acond;
        info->rnorm = rnorm; info->Arnorm = Arnorm; info->xnorm = xnorm;
    }
}

// c_i = dot(eta_i, xi_bar)
__global__ void eta_dots_kernel(const double* __restrict__ E, int64_t lde, int nobs, int K,
                                const double* __restrict__ x, double* __restrict__ c) {
    const int i = blockIdx.x;
    if (i >= K) return;
    double s = 0.0;
    for (int j = threadIdx.x; j < nobs; j += blockDim.x) s += E[(int64_t)i * lde + j] * x[j];
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s = ls_warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) {
        double r = (lane < (blockDim.x >> 5)) ? red[lane] : 0.0;
        r = ls_warp_sum(r);
        if (lane == 0) c[i] = r;
    }
}

// s = X*beta; for i: s += xis[i] * c_i   (unfused multiply-add, reference order src/lsqr.jl:57-61)
__global__ void pcga_update_kernel(const double* __restrict__ Z, int64_t ld, int64_t n, int K,
                                   const double* __restrict__ Xmean, double beta, const double* __restrict__ c,
                                   double* __restrict__ s) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double acc = __dmul_rn(Xmean[r], beta);
    const double* zr = Z + r * ld;
    for (int i = 0; i < K; ++i) acc = __dadd_rn(acc, __dmul_rn(zr[i], c[i]));
    s[r] = acc;
}

// P[:, i] = s + delta*xi_i (i < K), s + delta*X, s + delta*s, s   -- unfused, bit-identical to
// the host expression `s + delta * v` (src/lsqr.jl:39-43)
__global__ void paramstorun_kernel(const double* __restrict__ Z, int64_t ldz, int64_t n, int K,
                                   const double* __restrict__ s, const double* __restrict__ Xmean, double delta,
                                   double* __restrict__ P, int64_t ldp) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= n) return;
    const double sr = s[r];
    for (int c = threadIdx.x; c < K + 3; c += blockDim.x) {
        double v;
        if (c < K) v = __dadd_rn(sr, __dmul_rn(delta, Z[r * ldz + c]));
        else if (c == K) v = __dadd_rn(sr, __dmul_rn(delta, Xmean[r]));
        else if (c == K + 1) v = __dadd_rn(sr, __dmul_rn(delta, sr));
        else v = sr;
        P[r * ldp + c] = v;
    }
}

// Xt[j, c] = Rdiag[j] * S[c0 + c, j]   (TALL nobs x ncols), S column-major Nred x nobs
__global__ void scaled_transpose_kernel(const double* __restrict__ S, int64_t lds, int64_t nobs, int64_t c0,
                                        int64_t ncols, const double* __restrict__ Rdiag, double* __restrict__ Xt,
                                        int64_t ld) {
    __shared__ double tile[32][33];
    const int64_t j0 = (int64_t)blockIdx.x * 32, cc0 = (int64_t)blockIdx.y * 32;
    for (int jy = threadIdx.y; jy < 32; jy += blockDim.y) {
        const int64_t j = j0 + jy, c = cc0 + threadIdx.x;       // read S[c0+c, j]: contiguous in c
        tile[jy][threadIdx.x] = (j < nobs && c < ncols) ? S[j * lds + c0 + c] : 0.0;
    }
    __syncthreads();
    for (int jy = threadIdx.y; jy < 32; jy += blockDim.y) {
        const int64_t j = j0 + jy, c = cc0 + threadIdx.x;
        if (j < nobs && c < ncols) Xt[j * ld + c] = Rdiag[j] * tile[jy][threadIdx.x];
    }
}

// ---- pcgadirect: saddle-point matrix and pseudo-inverse solve (reference src/direct.jl:49-58)
// Columns [c0, c0+nc) of the top-left block of M (column-major, pitch ldm): HQH + R, where
// O (TALL nobs x nc) holds HQH[:, c0:c0+nc].
__global__ void saddle_block_kernel(const double* __restrict__ O, int64_t ldo, int nobs, int c0, int nc,
                                    const double* __restrict__ Rdiag, const double* __restrict__ Rdense,
                                    double* __restrict__ M, int64_t ldm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;     // row (fast index of the column-major M)
    const int c = blockIdx.y;
    if (i >= nobs || c >= nc) return;
    const int j = c0 + c;
    double v = O ? O[(int64_t)i * ldo + c] : 0.0;
    if (Rdiag) { if (i == j) v += Rdiag[i]; }
    else v += Rdense[(int64_t)j * nobs + i];
    M[(int64_t)j * ldm + i] = v;
}
// Border [.. HX; HX' 0] and the identity block below (rows m .. 2m-1) that accumulates V.
__global__ void saddle_border_kernel(const double* __restrict__ HX, int nobs, double* __restrict__ M, int64_t ldm) {
    const int m = nobs + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    M[(int64_t)i * ldm + nobs] = (i < nobs) ? HX[i] : 0.0;   // last row
    M[(int64_t)nobs * ldm + i] = (i < nobs) ? HX[i] : 0.0;   // last column
    M[(int64_t)i * ldm + m + i] = 1.0;                       // identity (the block was zeroed)
}
// After the Jacobi sweeps the top m rows of column j hold u_j * sigma_j and the bottom m rows
// v_j:  x = sum_{sigma_j > tol} v_j (u_j' b) / sigma_j,  tol = eps * m * sigma_max  (Julia's
// pinv default rtol = eps * min(size), atol = 0).  Single CTA; sh holds 2m doubles + LS_WARPS.
__global__ void __launch_bounds__(LS_THREADS)
pinv_apply_kernel(const double* __restrict__ M, int64_t ldm, int m, const double* __restrict__ b,
                  double* __restrict__ x, int* __restrict__ rank_out) {
    extern __shared__ double sh[];
    double* sig2 = sh;            // m
    double* coef = sh + m;        // m
    double* red = sh + 2 * m;     // LS_WARPS
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int j = warp; j < m; j += LS_WARPS) {
        const double* col = M + (int64_t)j * ldm;
        double a = 0.0, d = 0.0;
        for (int i = lane; i < m; i += 32) { const double v = col[i]; a += v * v; d += v * b[i]; }
        a = ls_warp_sum(a); d = ls_warp_sum(d);
        if (lane == 0) { sig2[j] = a; coef[j] = d; }
    }
    __syncthreads();
    double mx = 0.0;
    for (int j = threadIdx.x; j < m; j += LS_THREADS) mx = fmax(mx, sig2[j]);
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < LS_WARPS; ++w) mx = fmax(mx, red[w]);
    const double tol = 2.220446049250313e-16 * (double)m * sqrt(mx);
    __syncthreads();
    int kept = 0;
    for (int j = threadIdx.x; j < m; j += LS_THREADS) {
        const bool keep = sqrt(sig2[j]) > tol;
        coef[j] = keep ? coef[j] / sig2[j] : 0.0;
        kept += keep ? 1 : 0;
    }
    const double total = ls_block_sum((double)kept, red);     // (syncs: coef is complete afterwards)
    if (threadIdx.x == 0) *rank_out = (int)(total + 0.5);
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += LS_THREADS) {
        const double* vrow = M + m + i;
        double acc = 0.0;
        for (int j = 0; j < m; ++j) acc += vrow[(int64_t)j * ldm] * coef[j];
        x[i] = acc;
    }
}

struct DevMem {
    gsi_ctx* ctx; void* p; size_t bytes;
    DevMem(gsi_ctx* c, size_t b) : ctx(c), p(pool_alloc(c, b ? b : 8)), bytes(b ? b : 8) {}
    ~DevMem() { pool_free(ctx, p, bytes); }
    double* d() const { return static_cast<double*>(p); }
};

static void upload(gsi_ctx* ctx, double* dst, const double* src, size_t count) {
    GSI_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
}

struct PcgaDev {
    DevMem E, HX, R;
    PcgaOp op;
    PcgaDev(gsi_ctx* ctx, int64_t nobs, int64_t K, const double* Eh, int64_t lde, const double* HXh,
            const double* Rdiag, const double* Rdense, int64_t ldr)
        : E(ctx, (size_t)nobs * K * 8), HX(ctx, (size_t)nobs * 8),
          R(ctx, Rdiag ? (size_t)nobs * 8 : (size_t)nobs * nobs * 8) {
        GSI_REQUIRE(nobs >= 1 && K >= 0, GSI_ERR_INVALID_ARGUMENT, "pcga: nobs >= 1 and K >= 0 required");
        GSI_REQUIRE(Eh || K == 0, GSI_ERR_INVALID_ARGUMENT, "pcga: null etas");
        GSI_REQUIRE(HXh != nullptr, GSI_ERR_INVALID_ARGUMENT, "pcga: null HX");
        GSI_REQUIRE((Rdiag != nullptr) != (Rdense != nullptr), GSI_ERR_INVALID_ARGUMENT,
                    "pcga: exactly one of Rdiag / Rdense must be given");
        GSI_REQUIRE(lde >= nobs, GSI_ERR_INVALID_ARGUMENT, "pcga: lde < nobs");
        if (K > 0)
            GSI_CUDA(cudaMemcpy2DAsync(E.d(), nobs * 8, Eh, lde * 8, nobs * 8, K, cudaMemcpyHostToDevice, ctx->stream));
        upload(ctx, HX.d(), HXh, nobs);
        if (Rdiag) upload(ctx, R.d(), Rdiag, nobs);
        else {
            GSI_REQUIRE(ldr >= nobs, GSI_ERR_INVALID_ARGUMENT, "pcga: ldr < nobs");
            GSI_CUDA(cudaMemcpy2DAsync(R.d(), nobs * 8, Rdense, ldr * 8, nobs * 8, nobs, cudaMemcpyHostToDevice, ctx->stream));
        }
        op.E = E.d(); op.HX = HX.d();
        op.Rdiag = Rdiag ? R.d() : nullptr;
        op.Rdense = Rdiag ? nullptr : R.d();
        op.nobs = (int)nobs; op.K = (int)K; op.lde = nobs; op.ldr = nobs;
    }
};

}  // namespace gsi

using namespace gsi;

template <typename F>
static int32_t guarded(F&& f) {
    try { f(); return GSI_OK; }
    catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error(e.what()); return GSI_ERR_INVALID_ARGUMENT; }
    catch (...) { set_last_error("unknown error"); return GSI_ERR_INVALID_ARGUMENT; }
}

GSI_API int32_t gsi_pcga_lowrank_matvec(gsi_ctx* ctx, int64_t nobs, int64_t K, const double* E, int64_t lde,
                                        const double* HX, const double* Rdiag, const double* Rdense, int64_t ldr,
                                        const double* x, double* v) {
    return guarded([&] {
        GSI_REQUIRE(ctx && x && v, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        PcgaDev dev(ctx, nobs, K, E, lde, HX, Rdiag, Rdense, ldr);
        DevMem xv(ctx, (size_t)(nobs + 1) * 16);
        upload(ctx, xv.d(), x, nobs + 1);
        const size_t smem = (size_t)(K + 1 + LS_WARPS) * sizeof(double);
        pcga_matvec_kernel<<<1, LS_THREADS, smem, ctx->stream>>>(dev.op, xv.d(), xv.d() + nobs + 1);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        GSI_CUDA(cudaMemcpyAsync(v, xv.d() + nobs + 1, (size_t)(nobs + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

GSI_API int32_t gsi_pcga_lsqr_solve(gsi_ctx* ctx, int64_t nobs, int64_t K, const double* E, int64_t lde,
                                    const double* HX, const double* Rdiag, const double* Rdense, int64_t ldr,
                                    const double* b, double atol, double btol, double conlim, int64_t maxiter,
                                    double* x_out, int64_t* itn_out, int32_t* istop_out) {
    return guarded([&] {
        GSI_REQUIRE(ctx && b && x_out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        const double sqrt_eps = 1.4901161193847656e-08;
        if (atol <= 0.0) atol = sqrt_eps;                   // IterativeSolvers defaults
        if (btol <= 0.0) btol = sqrt_eps;
        if (conlim <= 0.0) conlim = 1.0 / sqrt_eps;
        if (maxiter <= 0) maxiter = nobs + 1;               // maximum(size(A))
        PcgaDev dev(ctx, nobs, K, E, lde, HX, Rdiag, Rdense, ldr);
        const int64_t m = nobs + 1;
        const size_t smem = (size_t)(5 * m + K + 1 + LS_WARPS) * sizeof(double);
        GSI_REQUIRE(smem <= 200 * 1024, GSI_ERR_UNSUPPORTED, "pcga lsqr: system too large for the single-CTA solver");
        DevMem bx(ctx, (size_t)(2 * m) * 8 + sizeof(LsqrOut));
        upload(ctx, bx.d(), b, m);
        LsqrOut* info = reinterpret_cast<LsqrOut*>(bx.d() + 2 * m);
        GSI_CUDA(cudaFuncSetAttribute(pcga_lsqr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        pcga_lsqr_kernel<<<1, LS_THREADS, smem, ctx->stream>>>(dev.op, bx.d(), atol, btol, conlim, (int)maxiter,
                                                               bx.d() + m, info);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        LsqrOut h;
        GSI_CUDA(cudaMemcpyAsync(x_out, bx.d() + m, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaMemcpyAsync(&h, info, sizeof(LsqrOut), cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
        if (itn_out) *itn_out = (int64_t)h.itn;
        if (istop_out) *istop_out = (int32_t)h.istop;
    });
}

GSI_API int32_t gsi_pcga_update(gsi_ctx* ctx, const gsi_buf* Zk, int64_t K, const double* Xmean, const double* E,
                                int64_t lde, int64_t nobs, const double* x, double* s_host) {
    return guarded([&] {
        GSI_REQUIRE(ctx && Zk && Xmean && x && s_host, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        GSI_REQUIRE(Zk->layout == GSI_LAYOUT_TALL && Zk->cols >= K, GSI_ERR_DIMENSION_MISMATCH, "pcga_update: Zk needs K columns");
        GSI_REQUIRE(lde >= nobs, GSI_ERR_INVALID_ARGUMENT, "pcga_update: lde < nobs");
        const int64_t n = Zk->rows;
        DevMem Ed(ctx, (size_t)nobs * (K > 0 ? K : 1) * 8), xd(ctx, (size_t)(nobs + 1) * 8), c(ctx, (size_t)(K + 1) * 8),
            Xd(ctx, (size_t)n * 8), sd(ctx, (size_t)n * 8);
        if (K > 0)
            GSI_CUDA(cudaMemcpy2DAsync(Ed.d(), nobs * 8, E, lde * 8, nobs * 8, K, cudaMemcpyHostToDevice, ctx->stream));
        upload(ctx, xd.d(), x, nobs + 1);
        upload(ctx, Xd.d(), Xmean, n);
        if (K > 0) eta_dots_kernel<<<(unsigned)K, 256, 0, ctx->stream>>>(Ed.d(), nobs, (int)nobs, (int)K, xd.d(), c.d());
        const double beta = x[nobs];
        pcga_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(Zk->d, Zk->ld, n, (int)K, Xd.d(), beta,
                                                                                  c.d(), sd.d());
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx, 2);
        GSI_CUDA(cudaMemcpyAsync(s_host, sd.d(), (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

GSI_API int32_t gsi_pcga_paramstorun(gsi_ctx* ctx, const gsi_buf* Zk, int64_t K, const double* s, const double* Xmean,
                                     double delta, gsi_buf* P_out) {
    return guarded([&] {
        GSI_REQUIRE(ctx && Zk && s && Xmean && P_out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        const int64_t n = Zk->rows;
        GSI_REQUIRE(Zk->layout == GSI_LAYOUT_TALL && Zk->cols >= K, GSI_ERR_DIMENSION_MISMATCH, "paramstorun: Zk needs K columns");
        GSI_REQUIRE(P_out->layout == GSI_LAYOUT_TALL && P_out->rows == n && P_out->cols == K + 3,
                    GSI_ERR_DIMENSION_MISMATCH, "paramstorun: P must be TALL n x (K+3)");
        DevMem sd(ctx, (size_t)n * 8), Xd(ctx, (size_t)n * 8);
        upload(ctx, sd.d(), s, n);
        upload(ctx, Xd.d(), Xmean, n);
        dim3 block(32, 8);
        paramstorun_kernel<<<(unsigned)((n + 7) / 8), block, 0, ctx->stream>>>(Zk->d, Zk->ld, n, (int)K, sd.d(), Xd.d(),
                                                                               delta, P_out->d, P_out->ld);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

GSI_API int32_t gsi_sketch_apply(gsi_ctx* ctx, gsi_buf* S, const gsi_buf* V, gsi_buf* out) {
    return guarded([&] {
        GSI_REQUIRE(ctx && S && V && out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        dense_apply(ctx, S, 0, V, out, 1.0);
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
    });
}

// O = S * diag(Rd) * S'[:, c0:c0+nc] for successive column chunks (S COLMAJOR r x c, Rd device c),
// each handed to `consume(O, c0, nc)` as a TALL r x nc buffer.
template <typename F>
static void scaled_outer_chunks(gsi_ctx* ctx, gsi_buf* S, const double* Rd, F&& consume) {
    const int64_t r = S->rows, c = S->cols;
    for (int64_t c0 = 0; c0 < r; c0 += kMaxCols) {
        const int64_t nc = (r - c0 < kMaxCols) ? r - c0 : kMaxCols;
        BufPtr Xt = make_buf(ctx, GSI_LAYOUT_TALL, c, nc);          // (diag(Rd) S')[:, c0:c0+nc]
        BufPtr O = make_buf(ctx, GSI_LAYOUT_TALL, r, nc);
        dim3 grid((unsigned)((c + 31) / 32), (unsigned)((nc + 31) / 32)), block(32, 8);
        scaled_transpose_kernel<<<grid, block, 0, ctx->stream>>>(S->d, S->ld, c, c0, nc, Rd, Xt->d, Xt->ld);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        dense_apply(ctx, S, 0, Xt.get(), O.get(), 1.0);
        consume(O.get(), c0, nc);
    }
}

GSI_API int32_t gsi_sketch_cov(gsi_ctx* ctx, gsi_buf* S, const double* Rdiag, double* out_host, int64_t ldo) {
    return guarded([&] {
        GSI_REQUIRE(ctx && S && Rdiag && out_host, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        GSI_REQUIRE(S->layout == GSI_LAYOUT_COLMAJOR, GSI_ERR_INVALID_ARGUMENT, "sketch_cov: S must be COLMAJOR");
        const int64_t Nred = S->rows, nobs = S->cols;
        GSI_REQUIRE(ldo >= Nred, GSI_ERR_INVALID_ARGUMENT, "sketch_cov: ldo < Nred");
        DevMem Rd(ctx, (size_t)nobs * 8);
        upload(ctx, Rd.d(), Rdiag, nobs);
        scaled_outer_chunks(ctx, S, Rd.d(), [&](gsi_buf* O, int64_t c0, int64_t) {
            tall_download(O, out_host + c0 * ldo, ldo, 0, Nred);
        });
    });
}

GSI_API int32_t gsi_pcga_direct_solve(gsi_ctx* ctx, int64_t nobs, int64_t K, const double* E, int64_t lde,
                                      const double* HX, const double* Rdiag, const double* Rdense, int64_t ldr,
                                      const double* b, double* x_out, int64_t* rank_out) {
    return guarded([&] {
        GSI_REQUIRE(ctx && b && x_out, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        GSI_REQUIRE(nobs >= 1 && nobs < 4096, GSI_ERR_UNSUPPORTED, "pcga direct solve: nobs must be in 1..4095");
        PcgaDev dev(ctx, nobs, K, E, lde, HX, Rdiag, Rdense, ldr);     // validates and uploads E, HX, R
        const int m = (int)nobs + 1;
        const int64_t ldm = 2 * (int64_t)m;
        DevMem M(ctx, (size_t)ldm * m * 8), bx(ctx, (size_t)2 * m * 8);
        GSI_CUDA(cudaMemsetAsync(M.d(), 0, (size_t)ldm * m * 8, ctx->stream));
        upload(ctx, bx.d(), b, m);
        const dim3 blk(128), grd_all((unsigned)((nobs + 127) / 128));
        auto place = [&](const gsi_buf* O, int64_t c0, int64_t nc) {
            const dim3 grd(grd_all.x, (unsigned)nc);
            saddle_block_kernel<<<grd, blk, 0, ctx->stream>>>(O ? O->d : nullptr, O ? O->ld : 0, (int)nobs, (int)c0,
                                                              (int)nc, dev.op.Rdiag, dev.op.Rdense, M.d(), ldm);
            GSI_CUDA(cudaGetLastError());
            count_launch(ctx);
        };
        if (K > 0) {
            // HQH = sum_i eta_i eta_i' = E E'  (the reference's ger! loop, direct.jl:49-53) as one
            // tensor-core product per 256-column chunk; E is zero-padded to >= 8 columns
            const int64_t Kp = K < 8 ? 8 : K;
            BufPtr Eb = make_buf(ctx, GSI_LAYOUT_COLMAJOR, nobs, Kp);
            GSI_CUDA(cudaMemcpy2DAsync(Eb->d, Eb->ld * 8, dev.op.E, nobs * 8, nobs * 8, K, cudaMemcpyDeviceToDevice,
                                       ctx->stream));
            DevMem ones(ctx, (size_t)Kp * 8);
            std::vector<double> one_h((size_t)Kp, 1.0);
            upload(ctx, ones.d(), one_h.data(), Kp);
            GSI_CUDA(cudaStreamSynchronize(ctx->stream));              // one_h leaves scope below
            scaled_outer_chunks(ctx, Eb.get(), ones.d(), place);
        } else {
            for (int64_t c0 = 0; c0 < nobs; c0 += kMaxCols)
                place(nullptr, c0, (nobs - c0 < kMaxCols) ? nobs - c0 : kMaxCols);
        }
        saddle_border_kernel<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(dev.op.HX, (int)nobs, M.d(), ldm);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        jacobi_sweeps(ctx, M.d(), ldm, m, 2 * m, m);
        const size_t smem = (size_t)(2 * m + LS_WARPS) * sizeof(double);
        GSI_CUDA(cudaFuncSetAttribute(pinv_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int* drank = ctx->dflags + 2;
        pinv_apply_kernel<<<1, LS_THREADS, smem, ctx->stream>>>(M.d(), ldm, m, bx.d(), bx.d() + m, drank);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        int hrank = 0;
        GSI_CUDA(cudaMemcpyAsync(x_out, bx.d() + m, (size_t)m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaMemcpyAsync(&hrank, drank, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        GSI_CUDA(cudaStreamSynchronize(ctx->stream));
        if (rank_out) *rank_out = hrank;
    });
}
