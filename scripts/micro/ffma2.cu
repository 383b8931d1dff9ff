// Microbenchmark: does the packed FP32 pipe instruction FFMA2 (PTX fma.rn.f32x2, sm_100+) double the FP32 rate of FFMA?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
// Each thread runs 8 independent dependency chains; 148 x 8 blocks of 256 threads, ITER x 8 instructions per thread.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;
__global__ void k_ffma(float* out, float a, float b)
{
    float x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b)
{
    unsigned long long x[8], A, B;
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    A = *(unsigned long long*)&av; B = *(unsigned long long*)&bv;
    for (int i = 0; i < 8; i++) { float2 v = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i); x[i] = *(unsigned long long*)&v; }
    for (int it = 0; it < ITER; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
    }
    float s = 0; for (int i = 0; i < 8; i++) { float2 v = *(float2*)&x[i]; s += v.x + v.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; rep++) {
        float ms1, ms2;
        cudaEventRecord(e0); k_ffma<<<148 * 8, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms1, e0, e1);
        cudaEventRecord(e0); k_ffma2<<<148 * 8, 256>>>(d, 1.0001f, 0.5f); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms2, e0, e1);
        const double n = 148.0 * 8 * 256 * ITER * 8;
        printf("FFMA  %.3f ms  %.1f Tflop/s (2 flops x %g instr)\nFFMA2 %.3f ms  %.1f Tflop/s (4 flops x %g instr)\n", ms1, 2 * n / ms1 / 1e9, n, ms2, 4 * n / ms2 / 1e9, n);
    }
    return 0;
}
