// Microbenchmark: DFMA / DMUL / LDS / MUFU.RCP64H dependent-issue latency and throughput vs ILP on sm_100a.
// nvcc -arch=sm_100a -O3 -o /tmp/fp64_latency tools/fp64_latency.cu && /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double* out, int iters, double a, double b, long long* cycles) {
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int CHAINS>
void run(int warps_per_sm, const char* name) {
    double* out; long long* cyc; long long h;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    dfma_chain<CHAINS><<<148, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaDeviceSynchronize();
    dfma_chain<CHAINS><<<148, warps_per_sm * 32>>>(out, iters, 1.0000001, 1e-9, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    double per_inst = (double)h / ((double)iters * CHAINS);
    printf("%s warps/SM=%2d chains=%d: %.2f cycles per DFMA per warp -> %.3f DFMA/cycle/SMSP\n", name, warps_per_sm, CHAINS,
           per_inst, (warps_per_sm / 4.0) / per_inst);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1>(4, "lat"); run<2>(4, "ilp"); run<4>(4, "ilp"); run<8>(4, "ilp");
    run<1>(8, "tlp"); run<1>(16, "tlp"); run<1>(32, "tlp"); run<2>(16, "mix"); run<4>(16, "mix");
    return 0;
}
