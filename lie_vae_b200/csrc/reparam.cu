// Fused SO(3) reparameterize + wrapped log-density, forward and backward (sm_100a; FP32 in production, FP64 instantiation).
//
// Replaces, with ONE kernel per direction, the ~760 ATen launches behind
//   N0reparameterize.nsample            reparameterize.py:137-141   v = eps * sigma
//   rodrigues                           lie_tools.py:56-64          R(v) = exp(hat v)
//   SO3reparameterize.nsample           reparameterize.py:269-273   z = mu @ R(v)
//   SO3reparameterize.log_posterior     reparameterize.py:233-263   wrapped density, k = -K..K
//   utils.logsumexp                     utils.py:4-26
//
// Closed form used for the density (cos(theta + 2 pi k) = cos(theta)):
//   a      = sum_i (u_i / sigma_i)^2 / 2,  u = v/theta
//   t_k    = -a * th_k^2 + log max(th_k^2, 1e-3),   th_k = theta + 2 pi k
//   log_q  = -sum_i log sigma_i - 1.5 log 2pi - log max(2 - 2cos theta, 1e-3) + LSE_k t_k
// with 2 - 2cos(theta) = 4 sin^2(theta/2) (no cancellation).  torch.clamp semantics in the
// backward: a clamped term contributes no gradient; the bound itself passes (>=).
//
// Layout: mu (B,3,3) = 9-float rows, sigma (B,3), eps (n,B,3), z (n,B,3,3), log_q (n,B);
// the sample index is flat over (n,B), mu/sigma broadcast over n (b = i mod B).  One thread
// per sample; each CTA stages its 256-sample tile through shared memory with 128-bit
// coalesced accesses (rows of 9 and 3 floats are odd strides -> conflict-free LDS), all
// arithmetic in registers, outputs written back through the same staging buffers.
// HBM-bound: 100 B/sample forward, 148 B/sample backward.
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

// ------------------------------------------------------------------ staging of broadcast rows
// rows [i0, i0+rows) of a (n,B,W) view of a (B,W) tensor: contiguous unless the tile wraps.
template <typename T, int W, int TILE>
__device__ __forceinline__ void stage_bcast(T* __restrict__ dst, const T* __restrict__ src,
                                            int64_t i0, int rows, int64_t B) {
    // n == 1 (the training case): rows are not broadcast and the emulated 64-bit modulo is skipped
    const int64_t b0 = i0 < B ? i0 : i0 % B;
    if (b0 + rows <= B) {
        if (rows == TILE) tile_g2s_full<T, TILE * W, TILE>(dst, src + b0 * W);
        else tile_g2s(dst, src + b0 * W, rows * W);
    } else {
        for (int idx = threadIdx.x; idx < rows * W; idx += blockDim.x) {
            const int r = idx / W, c = idx - r * W;
            dst[idx] = __ldg(src + ((i0 + r) % B) * W + c);
        }
    }
}

// ------------------------------------------------------------------ forward
// EULER: additionally emit the ZYZ Euler angles of z (group_matrix_to_eazyz, lie_tools.py:178-180 -- what
// VAE.decode feeds the action decoder, vae.py:182) from the registers that hold z; z itself is then optional.
template <typename T, int KT, bool EULER, int TILE>
__global__ void __launch_bounds__(TILE)
so3_reparam_fwd_kernel(const T* __restrict__ mu, const T* __restrict__ sigma, const T* __restrict__ eps,
                       T* __restrict__ z, T* __restrict__ angles, T* __restrict__ log_q, int64_t total,
                       int64_t B, int krt) {
    __shared__ __align__(16) T s_m[TILE * 9];   // mu in, z out (same row, same thread)
    __shared__ __align__(16) T s_s[TILE * 3];
    __shared__ __align__(16) T s_e[TILE * 3];   // eps in, Euler angles out
    const int64_t i0 = int64_t(blockIdx.x) * TILE;
    const int rows = int(min(int64_t(TILE), total - i0));
    stage_bcast<T, 9, TILE>(s_m, mu, i0, rows, B);
    stage_bcast<T, 3, TILE>(s_s, sigma, i0, rows, B);
    const bool full = rows == TILE;       // every CTA but the last: compile-time copy loops
    if (full) tile_g2s_full<T, TILE * 3, TILE>(s_e, eps + i0 * 3);
    else tile_g2s(s_e, eps + i0 * 3, rows * 3);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        T m[9], sg[3], v[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = s_m[t * 9 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = s_s[t * 3 + j]; v[j] = s_e[t * 3 + j] * sg[j]; }
        RodriguesCtx<T> k;
        rodrigues_ctx(v, k);
        T R[9], zr[9];
        axis_angle_matrix(k.u, k.s, k.w, R);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                zr[r * 3 + c] = Sc<T>::fma(m[r * 3], R[c], Sc<T>::fma(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
                s_m[t * 9 + r * 3 + c] = zr[r * 3 + c];
            }
        if (EULER) {
            T q[4], e[3];
            mat_to_quat_fwd(zr, q);
            quat_to_eazyz_fwd(q, e);
#pragma unroll
            for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = e[j];
        }
        if (log_q != nullptr) {
            const T q0 = k.u[0] / sg[0], q1 = k.u[1] / sg[1], q2 = k.u[2] / sg[2];
            const T a = T(0.5) * (q0 * q0 + q1 * q1 + q2 * q2);
            const T lse = winding_lse<T, KT>(k.theta, a, krt);
            const T den = Sc<T>::max(T(2) * k.w, T(RP_CLAMP));
            log_q[i0 + t] = lse - (Sc<T>::log(sg[0]) + Sc<T>::log(sg[1]) + Sc<T>::log(sg[2])) - T(RP_LOG_2PI_1P5) - Sc<T>::log(den);
        }
    }
    __syncthreads();
    if (full) {
        if (z != nullptr) tile_s2g_full<T, TILE * 9, TILE>(z + i0 * 9, s_m);
        if (EULER) tile_s2g_full<T, TILE * 3, TILE>(angles + i0 * 3, s_e);
    } else {
        if (z != nullptr) tile_s2g(z + i0 * 9, s_m, rows * 9);
        if (EULER) tile_s2g(angles + i0 * 3, s_e, rows * 3);
    }
}

// ------------------------------------------------------------------ backward
// per-sample gradients: g_mu (total,9), g_sigma (total,3); for n > 1 the caller sums over n
// (lv_sum_leading_f32).  EULER: the upstream gradient arrives (also) as g_angles and is pulled back
// through matrix -> quaternion -> Euler on the recomputed z.
template <typename T, int KT, bool EULER, int TILE>
__global__ void __launch_bounds__(TILE)
so3_reparam_bwd_kernel(const T* __restrict__ mu, const T* __restrict__ sigma, const T* __restrict__ eps,
                       const T* __restrict__ gz, const T* __restrict__ gangles, const T* __restrict__ glq,
                       T* __restrict__ gmu, T* __restrict__ gsigma, int64_t total, int64_t B, int krt) {
    __shared__ __align__(16) T s_m[TILE * 9];
    __shared__ __align__(16) T s_g[TILE * 9];   // gz in, g_mu out
    __shared__ __align__(16) T s_s[TILE * 3];
    __shared__ __align__(16) T s_e[TILE * 3];   // eps in, g_sigma out
    __shared__ __align__(16) T s_a[EULER ? TILE * 3 : 4];   // g_angles in
    const int64_t i0 = int64_t(blockIdx.x) * TILE;
    const int rows = int(min(int64_t(TILE), total - i0));
    stage_bcast<T, 9, TILE>(s_m, mu, i0, rows, B);
    stage_bcast<T, 3, TILE>(s_s, sigma, i0, rows, B);
    const bool full = rows == TILE;
    if (full) {
        tile_g2s_full<T, TILE * 3, TILE>(s_e, eps + i0 * 3);
        if (gz != nullptr) tile_g2s_full<T, TILE * 9, TILE>(s_g, gz + i0 * 9);
        if (EULER) tile_g2s_full<T, TILE * 3, TILE>(s_a, gangles + i0 * 3);
    } else {
        tile_g2s(s_e, eps + i0 * 3, rows * 3);
        if (gz != nullptr) tile_g2s(s_g, gz + i0 * 9, rows * 9);
        if (EULER) tile_g2s(s_a, gangles + i0 * 3, rows * 3);
    }
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        T m[9], G[9], sg[3], ep[3], v[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) { m[j] = s_m[t * 9 + j]; G[j] = gz != nullptr ? s_g[t * 9 + j] : T(0); }
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = s_s[t * 3 + j]; ep[j] = s_e[t * 3 + j]; v[j] = ep[j] * sg[j]; }
        RodriguesCtx<T> k;
        rodrigues_ctx(v, k);
        T R[9];
        axis_angle_matrix(k.u, k.s, k.w, R);
        if (EULER) {
            T zr[9], q[4], ge[3], gq[4], gze[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    zr[r * 3 + c] = Sc<T>::fma(m[r * 3], R[c], Sc<T>::fma(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
#pragma unroll
            for (int j = 0; j < 3; ++j) ge[j] = s_a[t * 3 + j];
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
                s_g[t * 9 + r * 3 + c] = Sc<T>::fma(G[r * 3], R[c * 3], Sc<T>::fma(G[r * 3 + 1], R[c * 3 + 1], G[r * 3 + 2] * R[c * 3 + 2]));
                gR[r * 3 + c] = Sc<T>::fma(m[r], G[c], Sc<T>::fma(m[3 + r], G[3 + c], m[6 + r] * G[6 + c]));
            }
        T gth, gu[3];
        rodrigues_bwd_theta_u(k, gR, &gth, gu);
        T gs_direct[3] = {T(0), T(0), T(0)};
        const T gl = glq != nullptr ? glq[i0 + t] : T(0);
        if (glq != nullptr) {
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
        for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = Sc<T>::fma(gv[j], ep[j], gs_direct[j]);
    }
    __syncthreads();
    if (full) {
        tile_s2g_full<T, TILE * 9, TILE>(gmu + i0 * 9, s_g);
        tile_s2g_full<T, TILE * 3, TILE>(gsigma + i0 * 3, s_e);
    } else {
        tile_s2g(gmu + i0 * 9, s_g, rows * 9);
        tile_s2g(gsigma + i0 * 3, s_e, rows * 3);
    }
}

}  // namespace lv

// ====================================================================== C ABI
static int reparam_check(const char* name, int64_t n, int64_t B, int k) {
    if (n < 0 || B < 0 || k < 0) { lv::set_error("%s: negative size", name); return LV_ERR_ARG; }
    if (k > 64) { lv::set_error("%s: k=%d winding terms unsupported (max 64)", name, k); return LV_ERR_UNSUPPORTED; }
    if ((n * B + lv::RP_TILE / 2 - 1) / (lv::RP_TILE / 2) > 0x7fffffffLL) { lv::set_error("%s: too many samples", name); return LV_ERR_ARG; }
    return LV_OK;
}

template <typename T, bool EULER>
static int reparam_fwd(const char* name, const T* mu, const T* sigma, const T* eps, T* z, T* angles,
                       T* log_q, int64_t n, int64_t B, int k, void* stream) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || !eps || (EULER ? !angles : !z)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int TILE = sizeof(T) == 8 ? lv::RP_TILE / 2 : lv::RP_TILE;
    const unsigned grid = unsigned((total + TILE - 1) / TILE);
    if constexpr (sizeof(T) == 8) lv::so3_reparam_fwd_kernel<T, 0, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    else if (k == 3) lv::so3_reparam_fwd_kernel<T, 3, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    else if (k == 10) lv::so3_reparam_fwd_kernel<T, 10, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    else lv::so3_reparam_fwd_kernel<T, 0, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    return lv::check_launch(name);
}

template <typename T, bool EULER>
static int reparam_bwd(const char* name, const T* mu, const T* sigma, const T* eps, const T* gz,
                       const T* gangles, const T* glq, T* gmu, T* gsigma, int64_t n, int64_t B, int k,
                       void* stream) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || !eps || !gmu || !gsigma || (EULER && !gangles)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int TILE = sizeof(T) == 8 ? lv::RP_TILE / 2 : lv::RP_TILE;
    const unsigned grid = unsigned((total + TILE - 1) / TILE);
    if constexpr (sizeof(T) == 8) lv::so3_reparam_bwd_kernel<T, 0, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    else if (k == 3) lv::so3_reparam_bwd_kernel<T, 3, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    else if (k == 10) lv::so3_reparam_bwd_kernel<T, 10, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    else lv::so3_reparam_bwd_kernel<T, 0, EULER, TILE><<<grid, TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    return lv::check_launch(name);
}

#define LV_REPARAM_ENTRY(SFX, T)                                                                                                 \
    extern "C" int lv_so3_reparam_fwd_##SFX(const T* mu, const T* sigma, const T* eps, T* z, T* log_q, int64_t n, int64_t B,     \
                                            int k, void* stream) {                                                               \
        return reparam_fwd<T, false>("so3_reparam_fwd", mu, sigma, eps, z, nullptr, log_q, n, B, k, stream);                     \
    }                                                                                                                            \
    extern "C" int lv_so3_reparam_bwd_##SFX(const T* mu, const T* sigma, const T* eps, const T* gz, const T* glq, T* gmu,        \
                                            T* gsigma, int64_t n, int64_t B, int k, void* stream) {                              \
        return reparam_bwd<T, false>("so3_reparam_bwd", mu, sigma, eps, gz, nullptr, glq, gmu, gsigma, n, B, k, stream);         \
    }                                                                                                                            \
    /* reparameterize fused with matrix -> ZYZ Euler (the pose the action decoder consumes); z is optional */                    \
    extern "C" int lv_so3_reparam_eazyz_fwd_##SFX(const T* mu, const T* sigma, const T* eps, T* z, T* angles, T* log_q,          \
                                                  int64_t n, int64_t B, int k, void* stream) {                                   \
        return reparam_fwd<T, true>("so3_reparam_eazyz_fwd", mu, sigma, eps, z, angles, log_q, n, B, k, stream);                 \
    }                                                                                                                            \
    extern "C" int lv_so3_reparam_eazyz_bwd_##SFX(const T* mu, const T* sigma, const T* eps, const T* gz, const T* gangles,      \
                                                  const T* glq, T* gmu, T* gsigma, int64_t n, int64_t B, int k, void* stream) {  \
        return reparam_bwd<T, true>("so3_reparam_eazyz_bwd", mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, stream);    \
    }
LV_REPARAM_ENTRY(f32, float)
LV_REPARAM_ENTRY(f64, double)
