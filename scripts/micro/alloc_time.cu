// micro-benchmark: cost of large allocations (cudaMalloc vs cudaMallocAsync), first-touch clear, free
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__global__ void k_clear(uint4* p, size_t n) { for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4); }
int main()
{
    cudaStream_t s; cudaStreamCreate(&s);
    cudaFree(0);
    for (int rep = 0; rep < 2; ++rep)
    for (size_t gb : {8ull, 34ull, 68ull}) {
        const size_t bytes = gb << 30;
        void* p = nullptr;
        double t0 = now(); cudaError_t e = cudaMalloc(&p, bytes); double t1 = now();
        k_clear<<<148 * 8, 256, 0, s>>>((uint4*)p, bytes / 16); cudaStreamSynchronize(s); double t2 = now();
        k_clear<<<148 * 8, 256, 0, s>>>((uint4*)p, bytes / 16); cudaStreamSynchronize(s); double t3 = now();
        cudaFree(p); double t4 = now();
        printf("cudaMalloc      %3zu GB: alloc %8.2f ms  first clear %8.2f ms  second clear %8.2f ms  free %8.2f ms (%s)\n", gb, t1 - t0, t2 - t1, t3 - t2, t4 - t3, cudaGetErrorString(e));
        t0 = now(); e = cudaMallocAsync(&p, bytes, s); cudaStreamSynchronize(s); t1 = now();
        k_clear<<<148 * 8, 256, 0, s>>>((uint4*)p, bytes / 16); cudaStreamSynchronize(s); t2 = now();
        k_clear<<<148 * 8, 256, 0, s>>>((uint4*)p, bytes / 16); cudaStreamSynchronize(s); t3 = now();
        cudaFreeAsync(p, s); cudaStreamSynchronize(s); t4 = now();
        printf("cudaMallocAsync %3zu GB: alloc %8.2f ms  first clear %8.2f ms  second clear %8.2f ms  free %8.2f ms (%s)\n", gb, t1 - t0, t2 - t1, t3 - t2, t4 - t3, cudaGetErrorString(e));
    }
    return 0;
}
