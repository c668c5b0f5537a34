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
    printf("\"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
