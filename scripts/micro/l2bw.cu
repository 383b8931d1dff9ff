// Microbenchmark: L2 read bandwidth of this GPU (BASELINE.md §2 asks for the measured L2 peak next to the L2 % figures of the
// traversal kernels).  Every thread streams 128-bit loads over a buffer that fits the L2 (working sets 16 ... 96 MB, 126 MB L2 on
// B200: the two dies each cache their own copy of remote lines, so the figure drops once 2 x the set exceeds the L2) and over one
// that does not (2 GB: HBM).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2bw l2bw.cu && ./l2bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void rd(const uint4* __restrict__ p, size_t n, int passes, unsigned* sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int k = 0; k < passes; k++)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p + i));   // .cg: cache in L2 only
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}
int main()
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t maxBytes = (size_t)2 << 30;
    uint4* d; unsigned* sink; cudaMalloc(&d, maxBytes); cudaMalloc(&sink, 4); cudaMemset(d, 1, maxBytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const size_t sizesMB[] = {16, 32, 48, 64, 96, 2048};
    for (size_t mb : sizesMB) {
        const size_t n = mb * (1 << 20) / 16;
        const int passes = mb >= 2048 ? 2 : 64;
        rd<<<sms * 8, 512>>>(d, n, 2, sink);                       // warm the L2
        float best = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0); rd<<<sms * 8, 512>>>(d, n, passes, sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        printf("working set %5zu MB: %8.1f GB/s (%s)\n", mb, (double)mb * (1 << 20) * passes / best / 1e6, mb >= 2048 ? "HBM" : "L2");
    }
    return 0;
}
