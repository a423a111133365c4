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

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pantea
