// Scalar device helpers shared by the kernels (double and float instantiations).
#pragma once
#include <cuda_runtime.h>

namespace pantea {

// explicitly rounded, never-contracted arithmetic for the neighbour predicate
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }

// single-shift minimum image (reference box.py:112-117)
template <typename T>
__device__ __forceinline__ T min_image(T dx, T box) {
    const T half = (T)0.5 * box;
    if (dx > half) dx = sub_rn(dx, box);
    if (dx < -half) dx = add_rn(dx, box);
    return dx;
}

// r = sqrt((dx^2 + dy^2) + dz^2); exactly 0 for the zero vector (reference distance.py:73-77)
__device__ __forceinline__ double norm3_rn(double x, double y, double z) {
    return sqrt(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
}
__device__ __forceinline__ float norm3_rn(float x, float y, float z) {
    return __fsqrt_rn(add_rn(add_rn(mul_rn(x, x), mul_rn(y, y)), mul_rn(z, z)));
}

template <typename T> __device__ __forceinline__ T t_exp(T x);
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float t_exp<float>(float x) { return expf(x); }
template <typename T> __device__ __forceinline__ T t_tanh(T x);
template <> __device__ __forceinline__ double t_tanh<double>(double x) { return tanh(x); }
template <> __device__ __forceinline__ float t_tanh<float>(float x) { return tanhf(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_pow(T x, T y);
template <> __device__ __forceinline__ double t_pow<double>(double x, double y) { return pow(x, y); }
template <> __device__ __forceinline__ float t_pow<float>(float x, float y) { return powf(x, y); }
template <typename T> __device__ __forceinline__ void t_sincos(T x, T* s, T* c);
template <> __device__ __forceinline__ void t_sincos<double>(double x, double* s, double* c) { sincos(x, s, c); }
template <> __device__ __forceinline__ void t_sincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <typename T> __device__ __forceinline__ T t_log1p(T x);
template <> __device__ __forceinline__ double t_log1p<double>(double x) { return log1p(x); }
template <> __device__ __forceinline__ float t_log1p<float>(float x) { return log1pf(x); }

constexpr double kPi = 3.14159265358979323846;
// ((e + 1/e) / (e - 1/e))^3, reference cutoff.py:82
constexpr double kTanhPre = 2.2637537952253504;

// cutoff function value and derivative, fc(r) = [r < rc] f(r)  (reference cutoff.py:67-110)
template <typename T>
__device__ __forceinline__ void cutoff_eval(int type, T r, T rc, T& fc, T& dfc) {
    fc = (T)0; dfc = (T)0;
    if (!(r < rc)) return;
    switch (type) {
        case PANTEA_CUT_HARD: fc = (T)1; break;
        case PANTEA_CUT_COS: {
            T s, c; t_sincos<T>((T)kPi * r / rc, &s, &c);
            fc = (T)0.5 * (c + (T)1); dfc = (T)-0.5 * ((T)kPi / rc) * s;
            break;
        }
        case PANTEA_CUT_TANHU:
        case PANTEA_CUT_TANH: {
            const T pre = type == PANTEA_CUT_TANH ? (T)kTanhPre : (T)1;
            const T t = t_tanh<T>((T)1 - r / rc);
            const T t2 = t * t;
            fc = pre * t2 * t; dfc = pre * ((T)-3 / rc) * t2 * ((T)1 - t2);
            break;
        }
        case PANTEA_CUT_EXP: {
            const T x = r / rc, om = (T)1 - x * x;
            const T f = t_exp<T>((T)1 - (T)1 / om);
            fc = f; dfc = -f * (T)2 * r / (rc * rc * om * om);
            break;
        }
        case PANTEA_CUT_POLY1: fc = ((T)2 * r - (T)3) * r * r + (T)1; dfc = (T)6 * r * r - (T)6 * r; break;
        case PANTEA_CUT_POLY2: {
            const T r2 = r * r, r3 = r2 * r;
            fc = (((T)15 - (T)6 * r) * r - (T)10) * r3 + (T)1;
            dfc = (T)-30 * r2 * r2 + (T)60 * r3 - (T)30 * r2;
            break;
        }
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// branch-free fast paths for the triplet loop (full double accuracy up to ~1 ulp; not IEEE-rounded)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double a) {  // 1/a for normal positive a: MUFU seed + 2 Newton steps
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
// FP32 mode (tolerance 1e-5): the hardware approximations (MUFU, ~1-2 ulp) are used directly
__device__ __forceinline__ float fast_rcp(float a) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}

__device__ __forceinline__ double fast_sqrt(double a) {  // sqrt(a) for normal positive a
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a * y, y, 1.0);          // 1 - a y^2
    y = fma(0.5 * y, e, y);
    e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    double s = a * y;
    return fma(0.5 * y, fma(-s, s, a), s);   // one correction step on the root itself
}
__device__ __forceinline__ float fast_sqrt(float a) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
// 1/sqrt(a) for normal positive a, ~2 ulp
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-a * y, y, 1.0);
    return fma(0.5 * y, e, y);
}
__device__ __forceinline__ float fast_rsqrt(float a) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
// triplet-loop variant: two Newton steps on the reciprocal root already reach ~2 ulp, the final correction is dropped;
// a == 0 yields NaN (the caller's comparison then rejects the lane)
__device__ __forceinline__ double fast_sqrt_loop(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-a * y, y, 1.0);
    y = fma(0.5 * y, e, y);
    return a * y;
}
__device__ __forceinline__ float fast_sqrt_loop(float a) { return fast_sqrt(a); }

// Taylor coefficients 1/n!, n = 0..12, in constant memory so that they fold into the FMA operands
static __constant__ double kExpC[13] = {
    1.0, 1.0, 0.5, 1.66666666666666666667e-01, 4.16666666666666666667e-02, 8.33333333333333333333e-03,
    1.38888888888888888889e-03, 1.98412698412698412698e-04, 2.48015873015873015873e-05, 2.75573192239858906526e-06,
    2.75573192239858906526e-07, 2.50521083854417187751e-08, 2.08767569878680989792e-09};
static __constant__ double kExpK[4] = {1.4426950408889634, 6755399441055744.0 /* 1.5 * 2^52 */,
                                       -6.93147180369123816490e-01, -1.90821492927058770002e-10};

// exp(y), branch-free: magic-number range reduction, degree-12 polynomial on |f| <= ln2/2 (Estrin scheme for
// instruction-level parallelism), exponent patched in with integer arithmetic.  CLAMP: arguments below -700 are
// clamped (result ~1e-304 instead of 0/denormal); without CLAMP the caller guarantees -700 < y < 700.
template <bool CLAMP>
__device__ __forceinline__ double fast_exp_t(double y) {
    const double yc = CLAMP ? fmax(y, -700.0) : y;
    const double t = fma(yc, kExpK[0], kExpK[1]);
    const int k = __double2loint(t);
    const double kf = t - kExpK[1];
    double f = fma(kf, kExpK[2], yc);
    f = fma(kf, kExpK[3], f);
    const double f2 = f * f, f4 = f2 * f2, f8 = f4 * f4;
    const double a0 = fma(kExpC[1], f, kExpC[0]), a1 = fma(kExpC[3], f, kExpC[2]), a2 = fma(kExpC[5], f, kExpC[4]),
                 a3 = fma(kExpC[7], f, kExpC[6]), a4 = fma(kExpC[9], f, kExpC[8]), a5 = fma(kExpC[11], f, kExpC[10]);
    const double b0 = fma(a1, f2, a0), b1 = fma(a3, f2, a2), b2 = fma(a5, f2, a4);
    const double d0 = fma(b1, f4, b0), d1 = fma(kExpC[12], f4, b2);
    const double p = fma(d1, f8, d0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}
// Table-driven exp for the triplet loop: exp(y) = 2^k * T[i] * p(f) with n = round(y * 64/ln2) = 64 k + i,
// f = y - n ln2/64 (|f| <= ln2/128) and a degree-5 polynomial; T[i] = 2^(i/64) lives in shared memory
// (`exp2_table_fill`).  ~10 FP64 instructions instead of ~18; CLAMP: arguments below -700 return 0.
static __constant__ double kExpT[6] = {92.332482616893657 /* 64/ln2 */, 6755399441055744.0 /* 1.5 * 2^52 */,
                                       -6.93147180369123816490e-01 / 64.0, -1.90821492927058770002e-10 / 64.0,
                                       1.66666666666666666667e-01, 4.16666666666666666667e-02};
__device__ __forceinline__ void exp2_table_fill(double* tab, int tid, int nthreads) {
    for (int i = tid; i < 64; i += nthreads) tab[i] = exp2((double)i * (1.0 / 64.0));
}
__device__ __forceinline__ void exp2_table_fill(float*, int, int) {}
template <bool CLAMP>
__device__ __forceinline__ double fast_exp_tab(double y, const double* __restrict__ tab) {
    const double t = fma(y, kExpT[0], kExpT[1]);
    const int n = __double2loint(t);
    const double nf = t - kExpT[1];
    double f = fma(nf, kExpT[2], y);
    f = fma(nf, kExpT[3], f);
    double p = fma(8.33333333333333333333e-03, f, kExpT[5]);
    p = fma(p, f, kExpT[4]);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    p *= tab[n & 63];
    const int hi = __double2hiint(p) + ((n >> 6) << 20), lo = __double2loint(p);
    // y < -700 (sign bit set, magnitude bits above those of 700.0): the result underflows, return 0 (integer test)
    const bool under = CLAMP && (unsigned)__double2hiint(y) > 0xC085E000u;
    return __hiloint2double(under ? 0 : hi, under ? 0 : lo);
}
template <bool CLAMP>
__device__ __forceinline__ float fast_exp_tab(float y, const float*) { return __expf(y); }
// tanh(y / 2), y in [0, ~40], through the table-driven exponential: 1 - 2 / (exp(y) + 1)
template <typename T>
__device__ __forceinline__ T fast_tanh_half_tab(T y, const T* __restrict__ tab) {
    return fma((T)-2, fast_rcp(fast_exp_tab<false>(y, tab) + (T)1), (T)1);
}
// tanh(x), x in [0, ~20], through the table-driven exponential
template <typename T>
__device__ __forceinline__ T fast_tanh_pos_tab(T x, const T* __restrict__ tab) {
    return (T)1 - (T)2 * fast_rcp(fast_exp_tab<false>((T)2 * x, tab) + (T)1);
}

__device__ __forceinline__ double fast_exp(double y) { return fast_exp_t<true>(y); }
__device__ __forceinline__ double fast_exp_small(double y) { return fast_exp_t<false>(y); }
__device__ __forceinline__ float fast_exp_small(float y) { return __expf(y); }
__device__ __forceinline__ float fast_exp(float y) { return __expf(y); }

__device__ __forceinline__ double cospi_t(double x) { return cospi(x); }
__device__ __forceinline__ float cospi_t(float x) { return cospif(x); }

// tanh(x) for x in [0, ~20]: 1 - 2/(exp(2x) + 1).  Absolute error ~1e-16 (relative accuracy degrades only where
// tanh -> 0, i.e. where the cutoff function tanh^3 is itself negligible)
template <typename T>
__device__ __forceinline__ T fast_tanh_pos(T x) {
    return (T)1 - (T)2 * fast_rcp(fast_exp_small((T)2 * x) + (T)1);
}

// cutoff value from the squared distance (third leg of G3), hot path
template <typename T>
__device__ __forceinline__ T cutoff_value_sq(int type, T r2, T rc, T inv_rc) {
    const T r = fast_sqrt(r2);
    if (!(r < rc)) return (T)0;
    switch (type) {
        case PANTEA_CUT_TANHU: { const T t = fast_tanh_pos<T>((T)1 - r * inv_rc); return t * t * t; }
        case PANTEA_CUT_TANH: { const T t = fast_tanh_pos<T>((T)1 - r * inv_rc); return (T)kTanhPre * t * t * t; }
        case PANTEA_CUT_HARD: return (T)1;
        case PANTEA_CUT_COS: return (T)0.5 * (cospi_t(r * inv_rc) + (T)1);
        case PANTEA_CUT_EXP: { const T x = r * inv_rc; return fast_exp((T)1 - fast_rcp((T)1 - x * x)); }
        case PANTEA_CUT_POLY1: return ((T)2 * r - (T)3) * r * r + (T)1;
        case PANTEA_CUT_POLY2: return (((T)15 - (T)6 * r) * r - (T)10) * r * r * r + (T)1;
        default: return (T)0;
    }
}

// value only (third leg of G3)
template <typename T>
__device__ __forceinline__ T cutoff_value(int type, T r, T rc) {
    if (!(r < rc)) return (T)0;
    switch (type) {
        case PANTEA_CUT_HARD: return (T)1;
        case PANTEA_CUT_COS: { T s, c; t_sincos<T>((T)kPi * r / rc, &s, &c); return (T)0.5 * (c + (T)1); }
        case PANTEA_CUT_TANHU: { const T t = t_tanh<T>((T)1 - r / rc); return t * t * t; }
        case PANTEA_CUT_TANH: { const T t = t_tanh<T>((T)1 - r / rc); return (T)kTanhPre * t * t * t; }
        case PANTEA_CUT_EXP: { const T x = r / rc; return t_exp<T>((T)1 - (T)1 / ((T)1 - x * x)); }
        case PANTEA_CUT_POLY1: return ((T)2 * r - (T)3) * r * r + (T)1;
        case PANTEA_CUT_POLY2: return (((T)15 - (T)6 * r) * r - (T)10) * r * r * r + (T)1;
        default: return (T)0;
    }
}

// activation value and derivative (reference activation.py:7-60; EXP is exp(-x))
template <typename T>
__device__ __forceinline__ void activation_eval(int act, T x, T& y, T& dy) {
    switch (act) {
        case PANTEA_ACT_TANH: { const T t = t_tanh<T>(x); y = t; dy = (T)1 - t * t; break; }
        case PANTEA_ACT_LOGISTIC: { const T s = (T)1 / ((T)1 + t_exp<T>(-x)); y = s; dy = s * ((T)1 - s); break; }
        case PANTEA_ACT_SOFTPLUS: {
            const T s = (T)1 / ((T)1 + t_exp<T>(-x));
            y = (x > (T)0 ? x : (T)0) + t_log1p<T>(t_exp<T>(-(x > (T)0 ? x : -x))); dy = s;
            break;
        }
        case PANTEA_ACT_RELU: y = x > (T)0 ? x : (T)0; dy = x > (T)0 ? (T)1 : (T)0; break;
        case PANTEA_ACT_GAUSSIAN: { const T g = t_exp<T>((T)-0.5 * x * x); y = g; dy = -x * g; break; }
        case PANTEA_ACT_COS: { T s, c; t_sincos<T>(x, &s, &c); y = c; dy = -s; break; }
        case PANTEA_ACT_EXP: { const T e = t_exp<T>(-x); y = e; dy = -e; break; }
        case PANTEA_ACT_HARMONIC: y = x * x; dy = (T)2 * x; break;
        default: y = x; dy = (T)1; break;
    }
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- asynchronous global -> shared copies (LDGSTS), tracked by cp.async groups ----
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

}  // namespace pantea
