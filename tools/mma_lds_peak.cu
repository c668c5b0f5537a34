// Ceiling of the kcov/dense inner loop: DMMA.8x8x4 fed by LDS.64 fragments exactly as in
// the GEMM kernels (16 x 56-column warp strip, X tile [32][220], A tile [64][36]), with no
// generation, no barriers, no global traffic.  Tells how much of the 37 TF/s DMMA issue peak
// survives shared-memory operand fetch at 8 / 12 / 16 warps per SM.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NBW, int GEN>
__global__ void __launch_bounds__(512, 1) mma_lds_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    constexpr int ld = 220, AP = 36;
    double* xs = sm;                // [32][220]
    double* as = sm + 32 * ld;      // [64][36]
    for (int i = threadIdx.x; i < 32 * ld + 64 * AP; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int rg = (warp >> 2) & 3, cg = (warp + rg) & 3, nb0 = cg * 7;
    double acc[2][NBW][2];
    for (int h = 0; h < 2; ++h) for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0; acc[h][nb][1] = 0; }
    const double* arow0 = as + (rg * 16 + g) * AP + t;
    const double* arow1 = arow0 + 8 * AP;
    double gsum = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            if (GEN && (ks & 1) == 0) {     // one Gaussian kernel value per 2 k-steps, as in kcov_gemm
                double d0 = arow0[ks] - 0.3, d1 = arow1[ks] - 0.1, d2 = arow0[ks + 1] - 0.2;
                double r2 = d0 * d0; r2 += d1 * d1; r2 += d2 * d2;
                gsum += exp(-0.5 * r2);
            }
            const double a0 = arow0[ks * 4], a1 = arow1[ks * 4];
            const double* xrow = xs + (ks * 4 + t) * ld + nb0 * 8 + g;
#pragma unroll
            for (int nb = 0; nb < NBW; ++nb) {
                const double b = xrow[nb * 8];
                dmma(acc[0][nb][0], acc[0][nb][1], a0, b);
                dmma(acc[1][nb][0], acc[1][nb][1], a1, b);
            }
        }
    }
    double s = gsum;
    for (int h = 0; h < 2; ++h) for (int nb = 0; nb < NBW; ++nb) s += acc[h][nb][0] + acc[h][nb][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}


__constant__ double kTab[64];
__device__ __forceinline__ double fast_exp_neg(double y, const double* tab) {
    const double L = 92.33248261689366, MAGIC = 6755399441055744.0;
    const double HI = 0x1.62e42fef80000p-7, LO = 0x1.1cf79abc9e3b4p-42;
    const double t = fma(y, -L, MAGIC);
    const int k = __double2loint(t);
    const double kf = t - MAGIC;
    double r = fma(kf, -HI, -y);
    r = fma(kf, -LO, r);
    double p = 1.0 / 120.0;
    p = fma(p, r, 1.0 / 24.0); p = fma(p, r, 1.0 / 6.0); p = fma(p, r, 0.5); p = fma(p, r, 1.0); p = fma(p, r, 1.0);
    const double v = tab[k & 63] * p;
    const int hi = __double2hiint(v) + ((k >> 6) << 20);
    return __hiloint2double(hi, __double2loint(v));
}

// MODE 0: all warps MMA only. MODE 1: all warps MMA + 1 fast-exp value per 2 k-steps (interleaved).
// MODE 2: warp-specialised: warps >= NMMA only generate (12 values per tile-iteration, ILP 4),
//         the others only MMA.
template <int MODE, int NMMA>
__global__ void __launch_bounds__(512, 1) spec_kernel(double* out, int iters) {
    extern __shared__ double sm[];
    constexpr int ld = 220, AP = 36, NBW = 7;
    double* xs = sm; double* as = sm + 32 * ld; double* tab = as + 64 * AP;
    for (int i = threadIdx.x; i < 32 * ld + 64 * AP; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
    if (threadIdx.x < 64) tab[threadIdx.x] = 1.0 + threadIdx.x / 64.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double s = 0.0;
    if (MODE == 2 && warp >= NMMA) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                double v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double d0 = as[(it + e + b) & 1023] - 0.3, d1 = as[(it + 2 * e + b + 7) & 1023] - 0.1, d2 = as[(e + 3 * b) & 1023] - 0.2;
                    double r2 = d0 * d0; r2 = fma(d1, d1, r2); r2 = fma(d2, d2, r2);
                    v[e] = fast_exp_neg(r2, tab);
                }
                s += (v[0] + v[1]) + (v[2] + v[3]);
            }
        }
    } else {
        const int rg = (warp >> 2) & 3, cg = (warp + rg) & 3, nb0 = cg * 7;
        double acc[2][NBW][2];
        for (int h = 0; h < 2; ++h) for (int nb = 0; nb < NBW; ++nb) { acc[h][nb][0] = 0; acc[h][nb][1] = 0; }
        const double* arow0 = as + (rg * 16 + g) * AP + t;
        const double* arow1 = arow0 + 8 * AP;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                if (MODE == 1 && (ks & 1) == 0) {
                    const double d0 = arow0[ks] - 0.3, d1 = arow1[ks] - 0.1, d2 = arow0[ks + 1] - 0.2;
                    double r2 = d0 * d0; r2 = fma(d1, d1, r2); r2 = fma(d2, d2, r2);
                    s += fast_exp_neg(r2, tab);
                }
                const double a0 = arow0[ks * 4], a1 = arow1[ks * 4];
                const double* xrow = xs + (ks * 4 + t) * ld + nb0 * 8 + g;
#pragma unroll
                for (int nb = 0; nb < NBW; ++nb) {
                    const double b = xrow[nb * 8];
                    dmma(acc[0][nb][0], acc[0][nb][1], a0, b);
                    dmma(acc[1][nb][0], acc[1][nb][1], a1, b);
                }
            }
        }
        for (int h = 0; h < 2; ++h) for (int nb = 0; nb < NBW; ++nb) s += acc[h][nb][0] + acc[h][nb][1];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NMMA>
static void run_spec(int sms, double* out, const char* name) {
    const int iters = 4000, threads = 512;
    const size_t smem = (32 * 220 + 64 * 36 + 64) * sizeof(double);
    cudaFuncSetAttribute(spec_kernel<MODE, NMMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    spec_kernel<MODE, NMMA><<<sms, threads, smem>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        spec_kernel<MODE, NMMA><<<sms, threads, smem>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const int mma_warps = (MODE == 2) ? NMMA : 16;
    const double flops = (double)sms * mma_warps * iters * 8.0 * 14.0 * 512.0;
    printf("\"%s_tflops\": %.2f, ", name, flops / best * 1e-9);
}

template <int GEN>
static void run(int sms, int threads, double* out) {
    const int iters = 4000;
    const size_t smem = (32 * 220 + 64 * 36) * sizeof(double);
    cudaFuncSetAttribute(mma_lds_kernel<7, GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mma_lds_kernel<7, GEN><<<sms, threads, smem>>>(out, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        mma_lds_kernel<7, GEN><<<sms, threads, smem>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = (double)sms * (threads / 32) * iters * 8.0 * 14.0 * 512.0;
    printf("\"gen%d_warps%d_tflops\": %.2f, ", GEN, threads / 32, flops / best * 1e-9);
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    double* out; cudaMalloc(&out, sizeof(double) * prop.multiProcessorCount * 512);
    printf("{");
    for (int th : {256, 384, 512}) { run<0>(prop.multiProcessorCount, th, out); run<1>(prop.multiProcessorCount, th, out); }
    run_spec<0, 16>(prop.multiProcessorCount, out, "spec_mma_only16");
    run_spec<1, 16>(prop.multiProcessorCount, out, "spec_interleaved_fastexp16");
    run_spec<2, 12>(prop.multiProcessorCount, out, "spec_12mma_4gen");
    printf("\"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
