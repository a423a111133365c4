// Throughput of the legacy warp-level tensor instructions on sm_100a (used to size the pair filter's distance tiles):
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench.bin tools/mma_bench.cu && tools/mma_bench.bin
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
    float c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = blockIdx.x, b1 = blockIdx.x + 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (KIND == 1)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(b0));
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND>
void run(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    float* out; cudaMalloc(&out, sizeof(float) * sms * blocks_per_sm * 256);
    const int iters = 4096;
    k<KIND><<<sms * blocks_per_sm, 256>>>(out, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<sms * blocks_per_sm, 256>>>(out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mmas_per_sm = (double)iters * 8 * 8 * blocks_per_sm;  // 8 warps x 8 per iteration
    printf("%-28s blocks/SM %d: %.3f ms, %.3f warp-MMA per clock per SM (at %d MHz)\n", name, blocks_per_sm, ms,
           mmas_per_sm / (ms * 1e-3 * clk * 1e3), clk / 1000);
    cudaFree(out);
}

int main() {
    for (int b = 1; b <= 2; ++b) {
        run<0>("m16n8k8 tf32", b);
        run<2>("m16n8k4 tf32", b);
        run<1>("m16n8k16 bf16", b);
    }
    return 0;
}
