// Rank-k update of a window of a TALL buffer on the FP64 tensor cores (k <= 32):
//     W[rows x ncols] += alpha * P[rows x k] * U[k x ncols]
// -- the trailing-matrix update of the blocked LU (A22 -= L21 U12, reference lu(Y),
// src/RandMatFact.jl:60,68,72) and of the blocked Householder QR / dorgqr (Y2 -= V W2,
// :57-58,75-76).  With k = 16 the update is a pure streaming read-modify-write of W
// (intensity 2 flop/byte), so it is built for memory-level parallelism, not for the tensor
// pipe: one warp owns a 16-row x 32-column strip, loads its W fragments straight into the
// DMMA accumulators (8 independent 16-byte loads per lane), takes the P fragments from
// global memory (the 8 rows x 32 bytes of a fragment load are whole sectors) and the U
// fragments from a shared-memory copy, and stores the strip back; 8 warps per CTA, several
// CTAs per SM, no pipeline state.  (Round 1 ran this through the big dense GEMM kernel: one
// 64-row tile per SM in flight, a 56 KB X-tile copy per tile, 150 us per update at C3.)
#include "common.cuh"
#include "ptx.cuh"

namespace gsi {

constexpr int RU_THREADS = 256;
constexpr int RU_WARPS = RU_THREADS / 32;
constexpr int RU_MAXK = 32;

// U: k x ncols in a TALL buffer (pitch ldu = 4 mod 8 doubles -> conflict-free B-fragment loads)
__global__ void __launch_bounds__(RU_THREADS, 3)
rank_update_kernel(const double* __restrict__ P, int64_t ldp, int64_t rows, int kdim, const double* __restrict__ U,
                   int64_t ldu, int ncols, double* __restrict__ W, int64_t ldw, double alpha) {
    extern __shared__ double s_u[];                          // [kpad][ldu]
    const int kpad = (kdim + 3) & ~3;
    for (int i = threadIdx.x; i < kpad * (int)ldu; i += RU_THREADS) {
        const int r = i / (int)ldu;
        s_u[i] = r < kdim ? U[i] : 0.0;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ncg = (ncols + 31) / 32;                       // 32-column groups
    const int64_t nrg = (rows + 15) / 16;                    // 16-row groups
    const int64_t units = nrg * ncg;
    const int nks = kpad / 4;
    for (int64_t u = (int64_t)blockIdx.x * RU_WARPS + warp; u < units; u += (int64_t)gridDim.x * RU_WARPS) {
        const int64_t rg = u / ncg;
        const int cg = (int)(u - rg * ncg);
        const int64_t r0 = rg * 16 + g, r1 = r0 + 8;
        const int c0 = cg * 32;
        const bool ok0 = r0 < rows, ok1 = r1 < rows;
        double acc[2][4][2];
        // W fragments -> accumulators (columns beyond ncols are neither read nor written)
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int c = c0 + nb * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool ok = h ? ok1 : ok0;
                const double* w = W + (h ? r1 : r0) * ldw + c;
                if (ok && c + 1 < ncols) {
                    const double2 v = *reinterpret_cast<const double2*>(w);
                    acc[h][nb][0] = v.x; acc[h][nb][1] = v.y;
                } else {
                    acc[h][nb][0] = (ok && c < ncols) ? w[0] : 0.0;
                    acc[h][nb][1] = 0.0;
                }
            }
        }
        for (int ks = 0; ks < nks; ++ks) {
            const int k = ks * 4 + t;
            const double a0 = (ok0 && k < kdim) ? alpha * P[r0 * ldp + k] : 0.0;
            const double a1 = (ok1 && k < kdim) ? alpha * P[r1 * ldp + k] : 0.0;
            const double* urow = s_u + (size_t)k * ldu + c0 + g;
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const double b = (c0 + nb * 8 + g < ncols) ? urow[nb * 8] : 0.0;
                dmma884(acc[0][nb][0], acc[0][nb][1], a0, b);
                dmma884(acc[1][nb][0], acc[1][nb][1], a1, b);
            }
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const int c = c0 + nb * 8 + 2 * t;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool ok = h ? ok1 : ok0;
                double* w = W + (h ? r1 : r0) * ldw + c;
                if (ok && c + 1 < ncols) *reinterpret_cast<double2*>(w) = make_double2(acc[h][nb][0], acc[h][nb][1]);
                else if (ok && c < ncols) w[0] = acc[h][nb][0];
            }
        }
    }
}

// W[rows x ncols] (pointer + pitch, a window of a TALL buffer) += alpha * P * X, where P is a
// rows x kdim window of a TALL buffer (pointer Pd, pitch ldp) and X is a small TALL buffer
// kdim x ncols.  alpha = +-1 scales the P fragments exactly.
void tall_window_update(gsi_ctx* ctx, const double* Pd, int64_t ldp, int64_t rows, int64_t kdim, const gsi_buf* X,
                        double* Wd, int64_t ldw, double alpha) {
    if (rows <= 0 || X->cols <= 0 || kdim <= 0) return;
    GSI_REQUIRE(X->layout == GSI_LAYOUT_TALL && X->rows == kdim, GSI_ERR_DIMENSION_MISMATCH, "window update: X rows");
    GSI_REQUIRE(kdim <= RU_MAXK, GSI_ERR_UNSUPPORTED, "window update: rank above 32");
    GSI_REQUIRE(((uintptr_t)Wd % 16 == 0) && (ldw % 2 == 0), GSI_ERR_INVALID_ARGUMENT,
                "window update: the W window must be 16-byte aligned");
    const int ncols = (int)X->cols;
    const int kpad = ((int)kdim + 3) & ~3;
    const size_t smem = (size_t)kpad * X->ld * sizeof(double);
    GSI_REQUIRE(smem <= 160 * 1024, GSI_ERR_UNSUPPORTED, "window update: rank x width exceeds the shared-memory copy of U");
    static bool attr_set = false;
    if (!attr_set) {
        GSI_CUDA(cudaFuncSetAttribute(rank_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    const int64_t units = ((rows + 15) / 16) * ((ncols + 31) / 32);
    int64_t grid = (units + RU_WARPS - 1) / RU_WARPS;
    const int64_t cap = (int64_t)ctx->num_sms * 6;
    if (grid > cap) grid = cap;
    rank_update_kernel<<<(unsigned)grid, RU_THREADS, smem, ctx->stream>>>(Pd, ldp, rows, (int)kdim, X->d, X->ld, ncols, Wd,
                                                                          ldw, alpha);
    GSI_CUDA(cudaGetLastError());
    count_launch(ctx);
}

}  // namespace gsi
