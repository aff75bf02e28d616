// Per-sample math of the fused SO(3) reparameterize kernels (shared by reparam.cu and head_reparam.cu).
// See reparam.cu for the formulas and the reference lines they replace.
#pragma once
#include "common.cuh"

namespace lv {

constexpr int RP_TILE = 256;          // float; the double instantiation uses 128-sample tiles (same smem footprint)
constexpr double RP_CLAMP = 1e-3;
constexpr double RP_TWO_PI = 6.283185307179586476925;
constexpr double RP_LOG_2PI_1P5 = 2.756815599614018102;   // 1.5 * log(2 pi)

// per-term exponential: float uses the MUFU base-2 exponential directly (the caller pre-scales the exponent by log2 e, so a
// winding costs FADD + FMUL + MUFU.EX2 + FFMA); double uses the library exp (scale 1).  The final log of the sum is full
// precision in both.
template <typename T> constexpr double RP_EXP_SCALE = 1.0;
template <> constexpr double RP_EXP_SCALE<float> = 1.442695040888963407;   // log2(e)
__device__ __forceinline__ float rp_exp2s(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ double rp_exp2s(double x) { return ::exp(x); }
// theta + 2 pi k with the product rounded separately (never contracted into an FMA): the reference point of the sum and
// the matching term of the loop are then bit-identical, so that term's exponent is exactly 0 -- and it is the rounding the
// reference's own tensor expression theta + 2 pi k performs (reparameterize.py:246-247).
__device__ __forceinline__ float winding_theta(float theta, float k) { return theta + __fmul_rn(float(RP_TWO_PI), k); }
__device__ __forceinline__ double winding_theta(double theta, double k) { return theta + __dmul_rn(RP_TWO_PI, k); }

// ------------------------------------------------------------------ in-kernel noise (Philox4x32-10, keyed by the sample index)
// The reference draws eps ~ N(0, 1) of shape (n, B, 3) from torch's global generator (reparameterize.py:137-141) and pins no
// stream, so any reproducible N(0, 1) source is admissible.  With eps == nullptr the kernels generate it: counter = global
// flat sample index (+ offset), key = seed, one Philox block per sample -> Box-Muller -> three normals.  The backward
// regenerates the same numbers, so eps never exists in memory (12 of the 60 input bytes per sample); lv_philox_normal_*
// materialises the identical stream for tests and for callers that want to inspect it.
__device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t key, uint32_t (&x)[4]) {
    uint32_t c0 = uint32_t(ctr), c1 = uint32_t(ctr >> 32), c2 = 0u, c3 = 0u;
    uint32_t k0 = uint32_t(key), k1 = uint32_t(key >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    x[0] = c0; x[1] = c1; x[2] = c2; x[3] = c3;
}
// three N(0, 1) samples of flat sample index `index`: u = (x + 0.5) 2^-32 in (0, 1); Box-Muller on (x0, x1) and (x2, x3)
template <typename T>
__device__ __forceinline__ void philox_normal3(uint64_t seed, uint64_t index, T (&ep)[3]) {
    uint32_t x[4];
    philox4x32_10(index, seed, x);
    const float u0 = (float(x[0] >> 8) + 0.5f) * 5.9604644775390625e-08f;     // 24 bits: exact in float, never 0 or 1
    const float u2 = (float(x[2] >> 8) + 0.5f) * 5.9604644775390625e-08f;
    const float a1 = float(x[1]) * 4.656612873077392578125e-10f;               // 2 u in [0, 2]: the angle 2 pi u as pi * a
    const float a3 = float(x[3]) * 4.656612873077392578125e-10f;
    const float r0 = sqrtf(-2.0f * logf(u0)), r2 = sqrtf(-2.0f * logf(u2));
    float s1, c1, s3, c3;
    sincospif(a1, &s1, &c1);
    sincospif(a3, &s3, &c3);
    ep[0] = T(r0 * c1);
    ep[1] = T(r0 * s1);
    ep[2] = T(r2 * c3);
    (void)s3;
}

// ------------------------------------------------------------------ wrapped log-density terms
// Term k of the winding sum is  t_k = -a x_k + log max(x_k, c),  x_k = (theta + 2 pi k)^2  (reparameterize.py:246-259), so
//   exp(t_k - ref) = max(x_k, c) * E_k,   E_k = exp(-a (x_k - x_ref)),
// and neither the per-term log nor (in the backward) the per-term 2/th_k divide is needed: one exponential per
// winding (the exponent is formed from the difference x_k - x_ref, so its rounding error is relative to the exponent, not to
// a x_k as in a per-term evaluation).  x_ref = x of the winding nearest to zero among k = -K..K (every exponent <= 0, the nearest term has E = 1, and
// the sum stays within [c, (2K+1) (theta + 2 pi K)^2]: no overflow or underflow).  LSE = -a x_ref + log(sum).
// KT > 0: winding count known at compile time (fully unrolled); KT == 0: runtime K.
template <typename T>
__device__ __forceinline__ T winding_ref(T theta, int K) {
    // nearest allowed winding: k* = clamp(round(-theta / 2 pi), -K, K)
    T ks = Sc<T>::rint(-theta * T(1.0 / RP_TWO_PI));
    ks = Sc<T>::max(T(-K), Sc<T>::min(T(K), ks));
    const T th = winding_theta(theta, ks);
    return th * th;
}

// returns sum_k max(x_k, c) E_k and the reference x_ref:  LSE_k t_k = -a x_ref + log(sum)
template <typename T, int KT>
__device__ __forceinline__ T winding_sum(T theta, T a, int krt, T* xref) {
    const int K = KT > 0 ? KT : krt;
    const T xr = winding_ref(theta, K);
    *xref = xr;
    const T a2 = -a * T(RP_EXP_SCALE<T>);         // exponent in the base rp_exp2s works in
    T s = T(0);
    if constexpr (KT > 0) {
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const T th = winding_theta(theta, T(i - KT));
            const T x = th * th;
            s = Sc<T>::fma(Sc<T>::max(x, T(RP_CLAMP)), rp_exp2s(a2 * (x - xr)), s);
        }
    } else {
        for (int k = -krt; k <= krt; ++k) {
            const T th = winding_theta(theta, T(k));
            const T x = th * th;
            s = Sc<T>::fma(Sc<T>::max(x, T(RP_CLAMP)), rp_exp2s(a2 * (x - xr)), s);
        }
    }
    return s;
}

// log(s0 s1 s2): one log of the product while no partial product can leave the normal range, else the three logs
template <typename T>
__device__ __forceinline__ T log_prod3(const T (&sg)[3]) {
    const T lo = Sc<T>::min(sg[0], Sc<T>::min(sg[1], sg[2])), hi = Sc<T>::max(sg[0], Sc<T>::max(sg[1], sg[2]));
    if (lo > T(1e-12) && hi < T(1e12)) return Sc<T>::log(sg[0] * sg[1] * sg[2]);
    return Sc<T>::log(sg[0]) + Sc<T>::log(sg[1]) + Sc<T>::log(sg[2]);
}

// softmax-weighted sums needed by the backward, w_k = max(x_k, c) E_k / sum:
//   d1 = sum_k w_k (-2 a th_k + [x_k >= c] 2/th_k) = sum_k E_k th_k ([x_k >= c] 2 - 2 a max(x_k, c)) / sum   (d LSE / d theta)
//   e2 = sum_k w_k x_k                                                                                         (-d LSE / d a)
template <typename T>
__device__ __forceinline__ void winding_grad_term(T th, T a, T a2, T xr, T& s, T& s1, T& s2) {
    const T x = th * th;
    const T e = rp_exp2s(a2 * (x - xr));
    const T xc = Sc<T>::max(x, T(RP_CLAMP));
    const T w = xc * e;
    const T dl = x >= T(RP_CLAMP) ? T(2) : T(0);
    s += w;
    s1 = Sc<T>::fma(e * th, Sc<T>::fma(T(-2) * a, xc, dl), s1);
    s2 = Sc<T>::fma(w, x, s2);
}

template <typename T, int KT>
__device__ __forceinline__ void winding_grad(T theta, T a, int krt, T* d1, T* e2) {
    const int K = KT > 0 ? KT : krt;
    const T xr = winding_ref(theta, K);
    const T a2 = -a * T(RP_EXP_SCALE<T>);
    T s = T(0), s1 = T(0), s2 = T(0);
    if constexpr (KT > 0) {
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) winding_grad_term(winding_theta(theta, T(i - KT)), a, a2, xr, s, s1, s2);
    } else {
        for (int k = -krt; k <= krt; ++k) winding_grad_term(winding_theta(theta, T(k)), a, a2, xr, s, s1, s2);
    }
    const T inv = T(1) / s;
    *d1 = s1 * inv;
    *e2 = s2 * inv;
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
        // log_q = LSE - sum_i log sigma_i - 1.5 log 2 pi - log den with LSE = -a x_ref + log(sum): two logs instead of five
        // (sum in [c, (2K+1) max x] and den in [c, 4]: the quotient stays far inside the normal range)
        T xr;
        const T sum = winding_sum<T, KT>(k.theta, a, krt, &xr);
        const T den = Sc<T>::max(T(2) * k.w, T(RP_CLAMP));
        *lq = Sc<T>::fma(-a, xr, Sc<T>::log(sum / den)) - log_prod3(sg) - T(RP_LOG_2PI_1P5);
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
