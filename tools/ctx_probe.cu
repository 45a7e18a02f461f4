// ctx_probe.cu -- where does process start-up go on a GPU box?  (tools only; cfg2's CLI wall is dominated by it)
//   nvcc -O2 -o build/ctx_probe tools/ctx_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
__global__ void k_nop(int *p) { if (p) *p = 1; }
int main()
{
    auto t0 = std::chrono::steady_clock::now();
    auto lap = [&](const char *what) { auto t = std::chrono::steady_clock::now(); printf("  %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(t - t0).count()); t0 = t; };
    cuInit(0); lap("cuInit");
    int n = 0; cudaGetDeviceCount(&n); lap("cudaGetDeviceCount");
    cudaSetDevice(0); cudaFree(0); lap("cudaSetDevice+cudaFree(0)");
    void *d = nullptr; cudaMalloc(&d, 24ull << 30); lap("cudaMalloc 24 GB");
    cudaMemset(d, 0, 1 << 20); cudaDeviceSynchronize(); lap("memset 1 MB + sync");
    k_nop<<<1, 32>>>((int *)d); cudaDeviceSynchronize(); lap("first kernel + sync");
    void *h = nullptr; cudaHostAlloc(&h, 64 << 20, cudaHostAllocDefault); lap("cudaHostAlloc 64 MB");
    cudaStream_t s; cudaStreamCreate(&s); lap("cudaStreamCreate");
    printf("  devices visible: %d\n", n);
    return 0;
}
