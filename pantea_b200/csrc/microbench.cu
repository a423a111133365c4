// Measurement helpers used by bench.py: FMA-pipe peak microbenchmark, L2 flush, work counters.
#include "internal.cuh"

namespace pantea {

// kChains independent FMA chains per thread: enough ILP to saturate the FP64 / FP32 pipe
template <typename T, int kChains>
__global__ void fma_peak_kernel(T* out, int iters, T a, T b) {
    T x[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) x[c] = (T)(threadIdx.x + c) * (T)1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) x[c] = x[c] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += x[c];
    if (s == (T)123.456) out[0] = s;  // never true: keeps the chains alive
}

__global__ void l2_flush_kernel(float4* buf, size_t n, float v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_float4(v, v, v, v);
}

}  // namespace pantea

using namespace pantea;

extern "C" {

// Launches the FMA microbenchmark; flops = 2 * kChains * iters * blocks * threads.  Time it with events.
int pantea_bench_fma(int32_t dtype, int32_t iters, int32_t blocks, int32_t threads, void* scratch, double* flops,
                     void* stream) {
    if (!scratch || iters < 1 || blocks < 1 || threads < 32) return fail(PANTEA_EINVAL, "pantea_bench_fma: bad argument");
    constexpr int kChains = 8;
    if (dtype == PANTEA_F64)
        fma_peak_kernel<double, kChains><<<blocks, threads, 0, (cudaStream_t)stream>>>((double*)scratch, iters, 1.0000001, 1e-9);
    else
        fma_peak_kernel<float, kChains><<<blocks, threads, 0, (cudaStream_t)stream>>>((float*)scratch, iters, 1.0000001f, 1e-9f);
    PANTEA_LAUNCH_CHECK();
    if (flops) *flops = 2.0 * kChains * (double)iters * (double)blocks * (double)threads;
    return PANTEA_OK;
}

// Overwrites `bytes` of scratch (> L2 size) so that the next kernel starts with a cold L2.
int pantea_l2_flush(void* scratch, int64_t bytes, void* stream) {
    if (!scratch || bytes < 16) return fail(PANTEA_EINVAL, "pantea_l2_flush: bad argument");
    l2_flush_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>((float4*)scratch, (size_t)(bytes / 16), 0.0f);
    PANTEA_LAUNCH_CHECK();
    return PANTEA_OK;
}

// counters: DEVICE uint64[4] (or NULL to disable): [0] neighbour pairs staged, [1] radial-SF evaluations,
// [2] triplet-SF evaluations, accumulated by every subsequent descriptor / energy launch of `ws`.
int pantea_workspace_set_counters(pantea_workspace* ws, void* counters) {
    if (!ws) return fail(PANTEA_EINVAL, "pantea_workspace_set_counters: NULL workspace");
    if (ws->counters != (unsigned long long*)counters) ++ws->arg_epoch;
    ws->counters = (unsigned long long*)counters;
    return PANTEA_OK;
}

}  // extern "C"
