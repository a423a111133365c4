// Throughput of packed single precision (FFMA2) against scalar FFMA on sm_100a: nvcc -arch=sm_100a -O3 tools/f32x2_bench.cu -o /tmp/f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_scalar(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_packed(float* out, int iters) {
    unsigned long long a[8];
    for (int i = 0; i < 8; ++i) { float x = threadIdx.x * 1e-3f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
    unsigned long long b, c;
    { float x = 1.0000001f, y = 1e-7f; asm("mov.b64 %0, {%1, %1};" : "=l"(b) : "f"(x)); asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(y)); }
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(b), "l"(c));
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i])); s += lo + hi; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
        float ms;
        k_scalar<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e0); k_scalar<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("scalar FFMA : %.3f ms, %.2f TFLOP/s, %.3f warp-instr/clk/SMSP\n", ms, 2.0 * 148 * 8 * 256 * 8.0 * iters / ms / 1e9, 148.0 * 8 * 8 * 8 * iters / (ms * 1e-3 * 1.965e9) / (148 * 4));
        k_packed<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e0); k_packed<<<148 * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("packed FFMA2: %.3f ms, %.2f TFLOP/s, %.3f warp-instr/clk/SMSP\n", ms, 4.0 * 148 * 8 * 256 * 8.0 * iters / ms / 1e9, 148.0 * 8 * 8 * 8 * iters / (ms * 1e-3 * 1.965e9) / (148 * 4));
    }
    return 0;
}
