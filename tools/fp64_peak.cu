// tools/fp64_peak.cu -- DFMA throughput microbenchmark: the fp64-pipe denominator for the rho/force kernels
// (MEASURED_PEAKS.json has no fp64 figure; SURVEY.md section 8d asks for one).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fma_per_s = (double)blocks * threads * iters * 64.0 / (ms * 1e-3);
        if (rep && fma_per_s > best) best = fma_per_s;
    }
    printf("{\"sms\": %d, \"dfma_per_s\": %.4e, \"fp64_tflops\": %.3f, \"dfma_per_clk_per_sm_at_1965MHz\": %.2f}\n", sms, best,
           2 * best / 1e12, best / sms / 1.965e9);
    return 0;
}
