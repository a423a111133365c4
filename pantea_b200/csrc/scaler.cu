// Per-feature statistics of a descriptor batch on the device: the reduction behind DescriptorScaler.fit /
// partial_fit (reference descriptors/scaler.py:250-283: jnp.mean / jnp.std / jnp.min / jnp.max over axis 0) and
// trainer.fit_scaler (potentials/nnp/trainer.py:68-88), the step that follows dataset preprocessing.
//
// Two passes over the [n_rows, n_cols] row-major matrix (mean first, then the centred second moment: the same
// two-pass definition as jnp.std), each a fixed-shape two-stage reduction -- the result does not depend on the launch
// or on atomics, so it is bitwise reproducible.  HBM-bound: 2 x n_rows x n_cols x sizeof(T) bytes.
#include "internal.cuh"
#include "math.cuh"

namespace pantea {

constexpr int kStatRowsPerBlock = 512;  // rows reduced by one block (8 row lanes x 64 iterations)
constexpr int kStatRowLanes = 8;

// partial[block][c][3] = (sum, min, max) of column c over the block's rows      (MOMENT == false)
// partial[block][c][0] = sum of (x - mean_c)^2                                   (MOMENT == true)
template <typename T, bool MOMENT>
__global__ void __launch_bounds__(32 * kStatRowLanes) stats_partial_kernel(const T* __restrict__ data, int64_t n_rows,
                                                                          int n_cols, int64_t ld,
                                                                          const double* __restrict__ mean,
                                                                          double* __restrict__ partial) {
    __shared__ double s_sum[kStatRowLanes][33], s_min[kStatRowLanes][33], s_max[kStatRowLanes][33];
    const int cl = threadIdx.x, rl = threadIdx.y;
    const int64_t r0 = (int64_t)blockIdx.x * kStatRowsPerBlock;
    const int64_t r1 = r0 + kStatRowsPerBlock < n_rows ? r0 + kStatRowsPerBlock : n_rows;
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
        const int c = c0 + cl;
        double sum = 0.0, mn = INFINITY, mx = -INFINITY;
        if (c < n_cols) {
            const double mu = MOMENT ? mean[c] : 0.0;
            for (int64_t r = r0 + rl; r < r1; r += kStatRowLanes) {
                const double x = (double)data[r * ld + c];
                if (MOMENT) { const double d = x - mu; sum += d * d; }
                else { sum += x; mn = fmin(mn, x); mx = fmax(mx, x); }
            }
        }
        s_sum[rl][cl] = sum; s_min[rl][cl] = mn; s_max[rl][cl] = mx;
        __syncthreads();
        if (rl == 0 && c < n_cols) {
            for (int q = 1; q < kStatRowLanes; ++q) {  // fixed order
                sum += s_sum[q][cl]; mn = fmin(mn, s_min[q][cl]); mx = fmax(mx, s_max[q][cl]);
            }
            double* o = partial + ((size_t)blockIdx.x * n_cols + c) * 3;
            o[0] = sum;
            if (!MOMENT) { o[1] = mn; o[2] = mx; }
        }
        __syncthreads();
    }
}

// stats[0][c] = mean, stats[1][c] = population sigma, stats[2][c] = min, stats[3][c] = max
template <bool MOMENT>
__global__ void stats_final_kernel(const double* __restrict__ partial, int n_blocks, int n_cols, int64_t n_rows,
                                   double* __restrict__ stats) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    double sum = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int b = 0; b < n_blocks; ++b) {  // fixed order
        const double* p = partial + ((size_t)b * n_cols + c) * 3;
        sum += p[0];
        if (!MOMENT) { mn = fmin(mn, p[1]); mx = fmax(mx, p[2]); }
    }
    if (MOMENT) stats[n_cols + c] = sqrt(sum / (double)n_rows);
    else { stats[c] = sum / (double)n_rows; stats[2 * n_cols + c] = mn; stats[3 * n_cols + c] = mx; }
}

template <typename T>
static int scaler_stats_typed(const T* data, int64_t n_rows, int n_cols, int64_t ld, double* stats, cudaStream_t st) {
    const int n_blocks = (int)((n_rows + kStatRowsPerBlock - 1) / kStatRowsPerBlock);
    double* partial = nullptr;
    PANTEA_CUDA_TRY(cudaMallocAsync((void**)&partial, sizeof(double) * 3 * (size_t)n_blocks * n_cols, st));
    const dim3 threads(32, kStatRowLanes);
    const int fin_blocks = (n_cols + 127) / 128;
    stats_partial_kernel<T, false><<<n_blocks, threads, 0, st>>>(data, n_rows, n_cols, ld, nullptr, partial);
    PANTEA_LAUNCH_CHECK();
    stats_final_kernel<false><<<fin_blocks, 128, 0, st>>>(partial, n_blocks, n_cols, n_rows, stats);
    PANTEA_LAUNCH_CHECK();
    stats_partial_kernel<T, true><<<n_blocks, threads, 0, st>>>(data, n_rows, n_cols, ld, stats, partial);
    PANTEA_LAUNCH_CHECK();
    stats_final_kernel<true><<<fin_blocks, 128, 0, st>>>(partial, n_blocks, n_cols, n_rows, stats);
    PANTEA_LAUNCH_CHECK();
    PANTEA_CUDA_TRY(cudaFreeAsync(partial, st));
    return PANTEA_OK;
}

}  // namespace pantea

using namespace pantea;

extern "C" int pantea_scaler_stats(const void* data, int64_t n_rows, int64_t n_cols, int64_t ld, int32_t dtype,
                                   double* stats, void* stream) {
    if (!data || !stats) return fail(PANTEA_EINVAL, "pantea_scaler_stats: NULL argument");
    if (n_rows < 1 || n_cols < 1 || ld < n_cols || n_cols > (1 << 20))
        return fail(PANTEA_EINVAL, "pantea_scaler_stats: need n_rows >= 1, 1 <= n_cols <= ld");
    if (dtype == PANTEA_F64) return scaler_stats_typed<double>((const double*)data, n_rows, (int)n_cols, ld, stats, (cudaStream_t)stream);
    if (dtype == PANTEA_F32) return scaler_stats_typed<float>((const float*)data, n_rows, (int)n_cols, ld, stats, (cudaStream_t)stream);
    return fail(PANTEA_EINVAL, "pantea_scaler_stats: dtype must be PANTEA_F32 or PANTEA_F64");
}
