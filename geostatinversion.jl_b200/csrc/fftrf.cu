// Batched FFTRF.powerlaw_structuredgrid on the device (reference
// src/FFTRF.jl:40-100, caller getxis(samplefield, ...) src/GeostatInversion.jl:29-38).
//
// Reference, per field:  sqrtS_f[a,b(,c)] = (fc1[b]^2 + fc2[a]^2 (+ fc3[c]^2))^(beta/4) on the DOUBLED
// grid of size (2N2, 2N1(, 2N3)) (first two dimensions swapped, :45-60), inf -> 0 (:65-67);
// result = sqrtS_f .* cis(2 pi phi), phi = randn(size(S)) (:74-80); k = ifft(result) (:92);
// finalk[j, i(, h)] = real(k[i, j(, h)]) over the first halves (:9-38); standardise to mean k0 and
// (corrected) standard deviation dk (:94-98).
//
// Here phi comes from the host (so `Random.seed!` keeps its meaning) for all fields at once, and
// the fields land as the columns of the n x N sample matrix that gsi_op_lowrankcov wraps -- the N
// host->device field uploads of the reference path (105 MB at C4) and the host FFTs disappear.
//
// The inverse transform is evaluated as a direct DFT per axis, restricted to the outputs that
// survive `reducek` (the first half of every axis): 6 N^3 complex multiply-adds per 2-D field
// instead of the 16 N^3 of a full DFT; at C4 (N = 256, 200 fields) that is 8e10 FP64 FMAs.  Twiddles
// come from an exact table w[m] = cis(2 pi m / L) (sincospi of a reduced integer argument), indexed
// by (o * k) mod L, so the result does not depend on the transform length being a power of two
// (the reference's tests use Ns = 25..50) and its rounding error is ~sqrt(L) eps.
#include "common.cuh"
#include "algos.h"

#define GSI_API extern "C" __attribute__((visibility("default")))

namespace gsi {

constexpr int FF_THREADS = 256;
constexpr int FF_MAXL = 2048;          // longest transform axis (doubled grid)

__device__ __forceinline__ double ff_fc(int idx, int N) {      // vcat(0:N, -(N-1):-1:-1)[idx]   (src/FFTRF.jl:88)
    return idx <= N ? (double)idx : (double)(idx - 2 * N);
}

// X[a + L2*(b + L1*c)] = sqrtS(a, b, c) * cis(2 pi phi)     (a: axis of fc2, b: axis of fc1, c: axis of fc3)
__global__ void ff_spectrum_kernel(const double* __restrict__ phi, int64_t ldphi, int N1, int N2, int N3, int dim,
                                   double beta, double2* __restrict__ X, int64_t fieldstride) {
    const int L1 = 2 * N1, L2 = 2 * N2, L3 = dim == 3 ? 2 * N3 : 1;
    const int64_t total = (int64_t)L1 * L2 * L3;
    const int f = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (int64_t)gridDim.x * blockDim.x) {
        const int a = (int)(j % L2), b = (int)((j / L2) % L1), c = (int)(j / ((int64_t)L1 * L2));
        const double f1 = ff_fc(b, N1), f2 = ff_fc(a, N2);
        double s = f1 * f1;                       // same accumulation order as :46-48 / :53-56
        s += f2 * f2;
        if (dim == 3) { const double f3 = ff_fc(c, N3); s += f3 * f3; }
        double v = pow(s, 0.25 * beta);
        if (isinf(v)) v = 0.0;                    // the zero frequency (:65-67)
        double sn, cs;
        sincospi(2.0 * phi[(int64_t)f * ldphi + j], &sn, &cs);
        X[(int64_t)f * fieldstride + j] = make_double2(v * cs, v * sn);
    }
}

// One axis of the inverse DFT (unnormalised), array viewed as [outer][L][inner] (inner fastest):
//   out[(o_out * nout + o) * inner + i] = sum_k in[(o_out * L + k) * inner + i] * w^(o k),  o < nout.
// A CTA stages TL lines (adjacent in `inner`, or whole contiguous lines when inner == 1) in shared memory.
__global__ void __launch_bounds__(FF_THREADS)
ff_axis_kernel(const double2* __restrict__ in, double2* __restrict__ out, int L, int nout, int64_t inner, int64_t outer,
               int TL) {
    extern __shared__ double2 ff_sm[];
    double2* w = ff_sm;                      // [L] twiddles
    double2* x = ff_sm + L;                  // [TL][L]
    for (int m = threadIdx.x; m < L; m += FF_THREADS) {
        double sn, cs;
        sincospi(2.0 * (double)m / (double)L, &sn, &cs);     // exact for the multiples of 1/4 turn
        w[m] = make_double2(cs, sn);
    }
    const int64_t ngroups_inner = (inner + TL - 1) / TL;
    const int64_t g = blockIdx.x;
    const int64_t o_out = g / ngroups_inner;
    const int64_t i0 = (g - o_out * ngroups_inner) * TL;
    const int tl = (int)((inner - i0 < TL) ? inner - i0 : TL);
    const double2* src = in + o_out * L * inner + i0;
    for (int idx = threadIdx.x; idx < L * tl; idx += FF_THREADS) {
        const int k = idx / tl, t = idx - k * tl;
        x[t * L + k] = src[(int64_t)k * inner + t];
    }
    __syncthreads();
    double2* dst = out + o_out * nout * inner + i0;
    for (int idx = threadIdx.x; idx < nout * tl; idx += FF_THREADS) {
        const int o = idx / tl, t = idx - o * tl;
        const double2* xt = x + t * L;
        double re = 0.0, im = 0.0;
        int m = 0;                             // (o * k) mod L
        for (int k = 0; k < L; ++k) {
            const double2 xv = xt[k], wv = w[m];
            re = fma(xv.x, wv.x, re); re = fma(-xv.y, wv.y, re);
            im = fma(xv.x, wv.y, im); im = fma(xv.y, wv.x, im);
            m += o;
            if (m >= L) m -= L;
        }
        dst[(int64_t)o * inner + t] = make_double2(re, im);
    }
    (void)outer;
}

__device__ double ff_block_sum(double v) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double r = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

// One CTA per field: finalk[j, i, h] = real(K[i, j, h]) / (L1 L2 L3) (K: [N3][N1][N2] with i < N2 fastest),
// mean, corrected standard deviation, out = dk (finalk - mean) / std + k0, written as column f of the
// n x N sample matrix in Julia order (j fastest, then i, then h).
__global__ void __launch_bounds__(1024)
ff_finalize_kernel(const double2* __restrict__ K, int64_t fieldstride, int N1, int N2, int N3, double scale, double k0,
                   double dk, double* __restrict__ out, int64_t ldo) {
    const int f = blockIdx.x;
    const double2* Kf = K + (int64_t)f * fieldstride;
    double* of = out + (int64_t)f * ldo;
    const int64_t n = (int64_t)N1 * N2 * N3;
    double s = 0.0;
    for (int64_t q = threadIdx.x; q < n; q += blockDim.x) {
        const int j = (int)(q % N1), i = (int)((q / N1) % N2), h = (int)(q / ((int64_t)N1 * N2));
        const double v = Kf[i + (int64_t)N2 * (j + (int64_t)N1 * h)].x * scale;
        of[q] = v;
        s += v;
    }
    const double mean = ff_block_sum(s) / (double)n;
    double ss = 0.0;
    for (int64_t q = threadIdx.x; q < n; q += blockDim.x) { const double d = of[q] - mean; ss += d * d; }
    const double sd = sqrt(ff_block_sum(ss) / (double)(n - 1));                      // Statistics.std: corrected
    for (int64_t q = threadIdx.x; q < n; q += blockDim.x) of[q] = dk * (of[q] - mean) / sd + k0;   // :96-98
}

void fftrf_powerlaw(gsi_ctx* ctx, int dim, const int64_t* Ns, double k0, double dk, double beta, const gsi_buf* phi,
                    gsi_buf* samples) {
    GSI_REQUIRE(dim == 2 || dim == 3, GSI_ERR_UNSUPPORTED, "fftrf: dimension must be 2 or 3 (src/FFTRF.jl:58)");
    const int N1 = (int)Ns[0], N2 = (int)Ns[1], N3 = dim == 3 ? (int)Ns[2] : 1;
    GSI_REQUIRE(N1 >= 1 && N2 >= 1 && N3 >= 1 && 2 * N1 <= FF_MAXL && 2 * N2 <= FF_MAXL && 2 * N3 <= FF_MAXL,
                GSI_ERR_UNSUPPORTED, "fftrf: grid sizes must be in 1..1024 per axis");
    const int L1 = 2 * N1, L2 = 2 * N2, L3 = dim == 3 ? 2 * N3 : 1;
    const int64_t big = (int64_t)L1 * L2 * L3, n = (int64_t)N1 * N2 * N3;
    GSI_REQUIRE(n >= 2, GSI_ERR_INVALID_ARGUMENT, "fftrf: a field needs at least 2 points");
    GSI_REQUIRE(phi->layout == GSI_LAYOUT_COLMAJOR && phi->rows == big, GSI_ERR_DIMENSION_MISMATCH,
                "fftrf: phi must be COLMAJOR prod(2 Ns) x nfields");
    const int64_t nf = phi->cols;
    GSI_REQUIRE(samples->layout == GSI_LAYOUT_COLMAJOR && samples->rows == n && samples->cols == nf,
                GSI_ERR_DIMENSION_MISMATCH, "fftrf: samples must be COLMAJOR prod(Ns) x nfields");
    cudaStream_t st = ctx->stream;
    // fields are processed in chunks that keep the two complex work arrays under ~1 GB
    int64_t chunk = ((int64_t)1 << 26) / big;                   // 2^26 complex = 1 GiB per array
    if (chunk < 1) chunk = 1;
    if (chunk > nf) chunk = nf;
    if (chunk > 65535) chunk = 65535;
    const size_t wbytes = (size_t)chunk * big * sizeof(double2);
    double2* A = static_cast<double2*>(pool_alloc(ctx, wbytes));
    double2* B = static_cast<double2*>(pool_alloc(ctx, wbytes));
    struct Guard { gsi_ctx* c; void* a; void* b; size_t n; ~Guard() { pool_free(c, a, n); pool_free(c, b, n); } } guard{ctx, A, B, wbytes};
    static bool attr_set = false;
    if (!attr_set) {
        GSI_CUDA(cudaFuncSetAttribute(ff_axis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    auto axis = [&](const double2* in, double2* out, int L, int nout, int64_t inner, int64_t outer) {
        int TL = (int)((180 * 1024 / sizeof(double2) - L) / L);     // lines staged per CTA
        if (TL > 16) TL = 16;
        if (TL > inner) TL = (int)inner;
        if (TL < 1) TL = 1;
        const int64_t groups = outer * ((inner + TL - 1) / TL);
        const size_t smem = (size_t)(L + (size_t)TL * L) * sizeof(double2);
        ff_axis_kernel<<<(unsigned)groups, FF_THREADS, smem, st>>>(in, out, L, nout, inner, outer, TL);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    };
    for (int64_t f0 = 0; f0 < nf; f0 += chunk) {
        const int64_t cf = (nf - f0 < chunk) ? nf - f0 : chunk;
        unsigned gx = (unsigned)((big + FF_THREADS - 1) / FF_THREADS);
        if (gx > 4096) gx = 4096;
        ff_spectrum_kernel<<<dim3(gx, (unsigned)cf), FF_THREADS, 0, st>>>(phi->d + f0 * phi->ld, phi->ld, N1, N2, N3, dim, beta,
                                                                         A, big);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
        // axis a (length L2, fastest): [cf * L3 * L1][L2][1] -> keep N2 outputs
        axis(A, B, L2, N2, 1, cf * (int64_t)L3 * L1);
        // axis b (length L1): [cf * L3][L1][N2] -> keep N1 outputs
        axis(B, A, L1, N1, N2, cf * (int64_t)L3);
        const double2* Kfin = A;
        int64_t fstride = (int64_t)L3 * N1 * N2;
        if (dim == 3) {
            // axis c (length L3): [cf][L3][N1 * N2] -> keep N3 outputs
            axis(A, B, L3, N3, (int64_t)N1 * N2, cf);
            Kfin = B;
            fstride = (int64_t)N3 * N1 * N2;
        }
        ff_finalize_kernel<<<(unsigned)cf, 1024, 0, st>>>(Kfin, fstride, N1, N2, N3, 1.0 / (double)big, k0, dk,
                                                          samples->d + f0 * samples->ld, samples->ld);
        GSI_CUDA(cudaGetLastError());
        count_launch(ctx);
    }
    GSI_CUDA(cudaStreamSynchronize(st));
}

}  // namespace gsi

using namespace gsi;

GSI_API int32_t gsi_fftrf_powerlaw(gsi_ctx* ctx, int32_t dim, const int64_t* Ns, double k0, double dk, double beta,
                                   const gsi_buf* phi, gsi_buf* samples) {
    try {
        GSI_REQUIRE(ctx && Ns && phi && samples, GSI_ERR_INVALID_ARGUMENT, "null argument");
        GSI_CUDA(cudaSetDevice(ctx->device));
        fftrf_powerlaw(ctx, dim, Ns, k0, dk, beta, phi, samples);
        return GSI_OK;
    } catch (const Error& e) { set_last_error(e.what()); return e.code; }
    catch (const std::exception& e) { set_last_error(e.what()); return GSI_ERR_INVALID_ARGUMENT; }
    catch (...) { set_last_error("unknown error"); return GSI_ERR_INVALID_ARGUMENT; }
}
