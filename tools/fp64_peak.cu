// Measures the FP64 issue-rate ceilings on this GPU (MEASURED_PEAKS.json has no FP64
// entry): register-resident DMMA.8x8x4 loop, DFMA loop, and a mixed loop.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC];
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// per iteration: NACC DMMAs + NF DFMAs (independent chains)
template <int NACC, int NF>
__global__ void mixed_kernel(double* out, int iters, double a0, double b0) {
    double c[NACC][2], f[NF > 0 ? NF : 1];
    for (int i = 0; i < NACC; ++i) { c[i][0] = 0; c[i][1] = 0; }
    for (int i = 0; i < NF; ++i) f[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    for (int i = 0; i < NF; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    double* out; cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
    const int iters = 20000;
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    for (int wps = 4; wps <= 16; wps *= 2) {           // warps per SM
        const int threads = 256, blocks = sms * wps * 32 / threads;
        float ms = time_ms([&] { dmma_kernel<16><<<blocks, threads>>>(out, iters, 1.0, 1.0); });
        double flops = (double)blocks * (threads / 32) * iters * 16.0 * 512.0;
        printf(", \"dmma_tflops_w%d\": %.2f", wps, flops / ms * 1e-9);
        ms = time_ms([&] { dfma_kernel<16><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
        flops = (double)blocks * threads * iters * 16.0 * 2.0;
        printf(", \"dfma_tflops_w%d\": %.2f", wps, flops / ms * 1e-9);
    }
    {   // mixed: 27 DMMA + 30 DFMA per iteration per warp, 8 warps per SM
        const int threads = 256, blocks = sms;
        float ms = time_ms([&] { mixed_kernel<27, 30><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
        double fl_mma = (double)blocks * (threads / 32) * iters * 27.0 * 512.0;
        double fl_fma = (double)blocks * threads * iters * 30.0 * 2.0;
        printf(", \"mixed_ms\": %.3f, \"mixed_dmma_tflops\": %.2f, \"mixed_dfma_tflops\": %.2f", ms, fl_mma / ms * 1e-9,
               fl_fma / ms * 1e-9);
        ms = time_ms([&] { mixed_kernel<27, 0><<<blocks, threads>>>(out, iters, 1.0, 1e-9); });
        printf(", \"dmma27_only_ms\": %.3f, \"dmma27_only_tflops\": %.2f", ms, fl_mma / ms * 1e-9);
    }
    printf("}\n");
    return 0;
}
