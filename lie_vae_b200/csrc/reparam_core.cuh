// Per-sample math of the fused SO(3) reparameterize kernels (shared by reparam.cu and head_reparam.cu).
// See reparam.cu for the formulas and the reference lines they replace.
#pragma once
#include "common.cuh"

namespace lv {

constexpr int RP_TILE = 256;          // float; the double instantiation uses 128-sample tiles (same smem footprint)
constexpr double RP_CLAMP = 1e-3;
constexpr double RP_TWO_PI = 6.283185307179586476925;
constexpr double RP_LOG_2PI_1P5 = 2.756815599614018102;   // 1.5 * log(2 pi)

// per-term log / exp / divide: fast intrinsics in float (one MUFU each; the final log of the LSE is full precision),
// the library functions in double
__device__ __forceinline__ float rp_log(float x) { return __logf(x); }
__device__ __forceinline__ float rp_exp(float x) { return __expf(x); }
__device__ __forceinline__ float rp_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double rp_log(double x) { return ::log(x); }
__device__ __forceinline__ double rp_exp(double x) { return ::exp(x); }
__device__ __forceinline__ double rp_div(double a, double b) { return a / b; }
template <typename T> __device__ __forceinline__ T rp_neg_inf();
template <> __device__ __forceinline__ float rp_neg_inf<float>() { return -INFINITY; }
template <> __device__ __forceinline__ double rp_neg_inf<double>() { return -HUGE_VAL; }

// ------------------------------------------------------------------ wrapped log-density terms
// KT > 0: winding count known at compile time (terms live in registers, one pass of logs);
// KT == 0: runtime K, two passes (max, then sum) recomputing the terms.
template <typename T, int KT>
__device__ __forceinline__ T winding_lse(T theta, T a, int krt) {
    if constexpr (KT > 0) {
        T t[2 * KT + 1];
        T m = rp_neg_inf<T>();
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const T th = theta + T(RP_TWO_PI) * T(i - KT);
            const T x = th * th;
            t[i] = Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP))));
            m = Sc<T>::max(m, t[i]);
        }
        T s = T(0);
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) s += rp_exp(t[i] - m);
        return m + Sc<T>::log(s);
    } else {
        T m = rp_neg_inf<T>();
        for (int k = -krt; k <= krt; ++k) {
            const T th = theta + T(RP_TWO_PI) * T(k);
            const T x = th * th;
            m = Sc<T>::max(m, Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP)))));
        }
        T s = T(0);
        for (int k = -krt; k <= krt; ++k) {
            const T th = theta + T(RP_TWO_PI) * T(k);
            const T x = th * th;
            s += rp_exp(Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP)))) - m);
        }
        return m + Sc<T>::log(s);
    }
}

// softmax-weighted sums needed by the backward:
//   d1 = sum_k w_k (-2 a th_k + [th_k^2 >= c] 2/th_k)    (d LSE / d theta)
//   e2 = sum_k w_k th_k^2                                 (-d LSE / d a)
template <typename T, int KT>
__device__ __forceinline__ void winding_grad(T theta, T a, int krt, T* d1, T* e2) {
    T m = rp_neg_inf<T>();
    if constexpr (KT > 0) {
        T t[2 * KT + 1];
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const T th = theta + T(RP_TWO_PI) * T(i - KT);
            const T x = th * th;
            t[i] = Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP))));
            m = Sc<T>::max(m, t[i]);
        }
        T s = T(0), s1 = T(0), s2 = T(0);
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const T th = theta + T(RP_TWO_PI) * T(i - KT);
            const T x = th * th;
            const T e = rp_exp(t[i] - m);
            const T dl = x >= T(RP_CLAMP) ? rp_div(T(2), th) : T(0);
            s += e;
            s1 = Sc<T>::fma(e, Sc<T>::fma(T(-2) * a, th, dl), s1);
            s2 = Sc<T>::fma(e, x, s2);
        }
        const T inv = T(1) / s;
        *d1 = s1 * inv;
        *e2 = s2 * inv;
    } else {
        for (int k = -krt; k <= krt; ++k) {
            const T th = theta + T(RP_TWO_PI) * T(k);
            const T x = th * th;
            m = Sc<T>::max(m, Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP)))));
        }
        T s = T(0), s1 = T(0), s2 = T(0);
        for (int k = -krt; k <= krt; ++k) {
            const T th = theta + T(RP_TWO_PI) * T(k);
            const T x = th * th;
            const T e = rp_exp(Sc<T>::fma(-a, x, rp_log(Sc<T>::max(x, T(RP_CLAMP)))) - m);
            const T dl = x >= T(RP_CLAMP) ? rp_div(T(2), th) : T(0);
            s += e;
            s1 = Sc<T>::fma(e, Sc<T>::fma(T(-2) * a, th, dl), s1);
            s2 = Sc<T>::fma(e, x, s2);
        }
        const T inv = T(1) / s;
        *d1 = s1 * inv;
        *e2 = s2 * inv;
    }
}

// ------------------------------------------------------------------ one sample, forward
// m = mean rotation (row-major), sg = sigma, ep = noise.  zr = m * exp(hat(ep * sg)); e = ZYZ Euler angles of zr (EULER);
// *lq = wrapped log-density (want_lq).
template <typename T, int KT, bool EULER>
__device__ __forceinline__ void reparam_sample_fwd(const T (&m)[9], const T (&sg)[3], const T (&ep)[3], int krt, bool want_lq,
                                                   T (&zr)[9], T (&e)[3], T* lq) {
    T v[3] = {ep[0] * sg[0], ep[1] * sg[1], ep[2] * sg[2]};
    RodriguesCtx<T> k;
    rodrigues_ctx(v, k);
    T R[9];
    axis_angle_matrix(k.u, k.s, k.w, R);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c)
            zr[r * 3 + c] = Sc<T>::fma(m[r * 3], R[c], Sc<T>::fma(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
    if (EULER) {
        T q[4];
        mat_to_quat_fwd(zr, q);
        quat_to_eazyz_fwd(q, e);
    }
    if (want_lq) {
        const T q0 = k.u[0] / sg[0], q1 = k.u[1] / sg[1], q2 = k.u[2] / sg[2];
        const T a = T(0.5) * (q0 * q0 + q1 * q1 + q2 * q2);
        const T lse = winding_lse<T, KT>(k.theta, a, krt);
        const T den = Sc<T>::max(T(2) * k.w, T(RP_CLAMP));
        *lq = lse - (Sc<T>::log(sg[0]) + Sc<T>::log(sg[1]) + Sc<T>::log(sg[2])) - T(RP_LOG_2PI_1P5) - Sc<T>::log(den);
    }
}

// ------------------------------------------------------------------ one sample, backward (recomputes the forward)
// G = upstream gradient of zr (zeros if none), ge = upstream gradient of the Euler angles (EULER), gl = upstream gradient
// of log_q (has_lq).  Outputs gm = d/d m, gsg = d/d sigma.
template <typename T, int KT, bool EULER>
__device__ __forceinline__ void reparam_sample_bwd(const T (&m)[9], const T (&sg)[3], const T (&ep)[3], T (&G)[9], const T (&ge)[3],
                                                   T gl, bool has_lq, int krt, T (&gm)[9], T (&gsg)[3]) {
    T v[3] = {ep[0] * sg[0], ep[1] * sg[1], ep[2] * sg[2]};
    RodriguesCtx<T> k;
    rodrigues_ctx(v, k);
    T R[9];
    axis_angle_matrix(k.u, k.s, k.w, R);
    if (EULER) {
        T zr[9], q[4], gq[4], gze[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                zr[r * 3 + c] = Sc<T>::fma(m[r * 3], R[c], Sc<T>::fma(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
        mat_to_quat_fwd(zr, q);
        quat_to_eazyz_bwd(q, ge, gq);
        mat_to_quat_bwd(zr, gq, gze);
#pragma unroll
        for (int j = 0; j < 9; ++j) G[j] += gze[j];
    }
    // z = mu R:  g_mu = gz R^T,  g_R = mu^T gz
    T gR[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            gm[r * 3 + c] = Sc<T>::fma(G[r * 3], R[c * 3], Sc<T>::fma(G[r * 3 + 1], R[c * 3 + 1], G[r * 3 + 2] * R[c * 3 + 2]));
            gR[r * 3 + c] = Sc<T>::fma(m[r], G[c], Sc<T>::fma(m[3 + r], G[3 + c], m[6 + r] * G[6 + c]));
        }
    T gth, gu[3];
    rodrigues_bwd_theta_u(k, gR, &gth, gu);
    T gs_direct[3] = {T(0), T(0), T(0)};
    if (has_lq) {
        const T is0 = T(1) / sg[0], is1 = T(1) / sg[1], is2 = T(1) / sg[2];
        const T q0 = k.u[0] * is0, q1 = k.u[1] * is1, q2 = k.u[2] * is2;
        const T a = T(0.5) * (q0 * q0 + q1 * q1 + q2 * q2);
        T d1, e2;
        winding_grad<T, KT>(k.theta, a, krt, &d1, &e2);
        // - d/dtheta log max(2w, c), 2w = 2 - 2cos: (2 sin)/(2w) = s/w where not clamped
        const T dden = (T(2) * k.w >= T(RP_CLAMP)) ? k.s / k.w : T(0);
        gth = Sc<T>::fma(gl, d1 - dden, gth);
        // d log_q / d a = -e2 ; d a / d u_i = u_i / sigma_i^2 ; d a / d sigma_i = -u_i^2 / sigma_i^3
        gu[0] = Sc<T>::fma(-gl * e2, q0 * is0, gu[0]);
        gu[1] = Sc<T>::fma(-gl * e2, q1 * is1, gu[1]);
        gu[2] = Sc<T>::fma(-gl * e2, q2 * is2, gu[2]);
        gs_direct[0] = gl * (e2 * q0 * q0 - T(1)) * is0;
        gs_direct[1] = gl * (e2 * q1 * q1 - T(1)) * is1;
        gs_direct[2] = gl * (e2 * q2 * q2 - T(1)) * is2;
    }
    T gv[3];
    theta_u_to_v(k, gth, gu, gv);
#pragma unroll
    for (int j = 0; j < 3; ++j) gsg[j] = Sc<T>::fma(gv[j], ep[j], gs_direct[j]);
}

}  // namespace lv
