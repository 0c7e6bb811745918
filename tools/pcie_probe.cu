// tools/pcie_probe.cu -- dev-only measurement: how fast do parts of a 104-byte AoS record array cross PCIe?
// (decides how misa_b200_step_host_fields moves masked fields). nvcc -O3 -arch=sm_100a -o tools/pcie_probe tools/pcie_probe.cu
//   A  contiguous D2H / H2D of the whole array (the full-record contract), and both directions at once
//   B  cudaMemcpy2DAsync: width 48 B (x|v) at pitch 104 B, packed on the device side
//   C  kernel storing / loading the 48 bytes straight to / from the mapped pinned host array
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_store_words(unsigned long long *host, const unsigned long long *packed, long long n, int w0, int nw) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * nw; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / nw;
        const int w = (int)(t - r * nw);
        host[r * 13 + w0 + w] = packed[t];
    }
}
__global__ void k_load_words(const unsigned long long *host, unsigned long long *packed, long long n, int w0, int nw) {
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n * nw; t += (long long)gridDim.x * blockDim.x) {
        const long long r = t / nw;
        const int w = (int)(t - r * nw);
        packed[t] = host[r * 13 + w0 + w];
    }
}
static float timed(cudaStream_t s, cudaEvent_t a, cudaEvent_t b) { CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 2000000;
    const size_t bytes = (size_t)n * 104;
    void *h = nullptr, *h2 = nullptr, *d = nullptr, *d2 = nullptr, *dp = nullptr, *hd = nullptr;
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped)); CK(cudaHostAlloc(&h2, bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&d2, bytes)); CK(cudaMalloc(&dp, (size_t)n * 72));
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    cudaEvent_t a, b, c2; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreate(&c2));
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaEventRecord(a, s1)); CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s1)); CK(cudaEventRecord(b, s1));
        float t = timed(s1, a, b); printf("A  D2H contiguous %zu MB: %.3f ms  %.1f GB/s\n", bytes >> 20, t, bytes / t / 1e6);
        CK(cudaEventRecord(a, s1)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1)); CK(cudaEventRecord(b, s1));
        t = timed(s1, a, b); printf("A  H2D contiguous: %.3f ms  %.1f GB/s\n", t, bytes / t / 1e6);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a, s1)); CK(cudaStreamWaitEvent(s2, a, 0));
        CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s1)); CK(cudaMemcpyAsync(d2, h2, bytes, cudaMemcpyHostToDevice, s2));
        CK(cudaEventRecord(c2, s2)); CK(cudaStreamWaitEvent(s1, c2, 0)); CK(cudaEventRecord(b, s1));
        t = timed(s1, a, b); printf("A  duplex D2H + H2D: %.3f ms  %.1f GB/s each way\n", t, bytes / t / 1e6);
        for (int nw = 6; nw <= 9; nw += 3) {
            const size_t payload = (size_t)n * nw * 8;
            CK(cudaEventRecord(a, s1));
            CK(cudaMemcpy2DAsync((char *)h + 16, 104, dp, nw * 8, nw * 8, n, cudaMemcpyDeviceToHost, s1));
            CK(cudaEventRecord(b, s1));
            t = timed(s1, a, b); printf("B  D2H 2-D copy width %d pitch 104: %.3f ms  %.1f GB/s payload\n", nw * 8, t, payload / t / 1e6);
            CK(cudaEventRecord(a, s1));
            CK(cudaMemcpy2DAsync(dp, nw * 8, (char *)h + 16, 104, nw * 8, n, cudaMemcpyHostToDevice, s1));
            CK(cudaEventRecord(b, s1));
            t = timed(s1, a, b); printf("B  H2D 2-D copy width %d pitch 104: %.3f ms  %.1f GB/s payload\n", nw * 8, t, payload / t / 1e6);
            CK(cudaEventRecord(a, s1));
            k_store_words<<<148 * 8, 256, 0, s1>>>((unsigned long long *)hd, (const unsigned long long *)dp, n, 2, nw);
            CK(cudaEventRecord(b, s1));
            t = timed(s1, a, b); printf("C  kernel stores of %d B per record into mapped host memory: %.3f ms  %.1f GB/s payload\n", nw * 8, t, payload / t / 1e6);
            CK(cudaEventRecord(a, s1));
            k_load_words<<<148 * 8, 256, 0, s1>>>((const unsigned long long *)hd, (unsigned long long *)dp, n, 2, nw);
            CK(cudaEventRecord(b, s1));
            t = timed(s1, a, b); printf("C  kernel loads of %d B per record from mapped host memory: %.3f ms  %.1f GB/s payload\n", nw * 8, t, payload / t / 1e6);
        }
    }
    return 0;
}
