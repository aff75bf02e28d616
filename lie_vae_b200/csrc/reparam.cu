// Fused SO(3) reparameterize + wrapped log-density, forward and backward (sm_100a, FP32).
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

constexpr int RP_TILE = 256;
constexpr float RP_CLAMP = 1e-3f;
constexpr float RP_TWO_PI = 6.283185307179586f;
constexpr float RP_LOG_2PI_1P5 = 2.756815599614018f;   // 1.5 * log(2 pi)

// ------------------------------------------------------------------ wrapped log-density terms
// KT > 0: winding count known at compile time (terms live in registers, one pass of logs);
// KT == 0: runtime K, two passes (max, then sum) recomputing the terms.
template <int KT>
__device__ __forceinline__ float winding_lse(float theta, float a, int krt) {
    if constexpr (KT > 0) {
        float t[2 * KT + 1];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const float th = theta + RP_TWO_PI * float(i - KT);
            const float x = th * th;
            t[i] = fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP)));
            m = fmaxf(m, t[i]);
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) s += __expf(t[i] - m);
        return m + logf(s);
    } else {
        float m = -INFINITY;
        for (int k = -krt; k <= krt; ++k) {
            const float th = theta + RP_TWO_PI * float(k);
            const float x = th * th;
            m = fmaxf(m, fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP))));
        }
        float s = 0.f;
        for (int k = -krt; k <= krt; ++k) {
            const float th = theta + RP_TWO_PI * float(k);
            const float x = th * th;
            s += __expf(fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP))) - m);
        }
        return m + logf(s);
    }
}

// softmax-weighted sums needed by the backward:
//   d1 = sum_k w_k (-2 a th_k + [th_k^2 >= c] 2/th_k)    (d LSE / d theta)
//   e2 = sum_k w_k th_k^2                                 (-d LSE / d a)
template <int KT>
__device__ __forceinline__ void winding_grad(float theta, float a, int krt, float* d1, float* e2) {
    float m = -INFINITY;
    if constexpr (KT > 0) {
        float t[2 * KT + 1];
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const float th = theta + RP_TWO_PI * float(i - KT);
            const float x = th * th;
            t[i] = fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP)));
            m = fmaxf(m, t[i]);
        }
        float s = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 2 * KT + 1; ++i) {
            const float th = theta + RP_TWO_PI * float(i - KT);
            const float x = th * th;
            const float e = __expf(t[i] - m);
            const float dl = x >= RP_CLAMP ? __fdividef(2.f, th) : 0.f;
            s += e;
            s1 = fmaf(e, fmaf(-2.f * a, th, dl), s1);
            s2 = fmaf(e, x, s2);
        }
        const float inv = 1.f / s;
        *d1 = s1 * inv;
        *e2 = s2 * inv;
    } else {
        for (int k = -krt; k <= krt; ++k) {
            const float th = theta + RP_TWO_PI * float(k);
            const float x = th * th;
            m = fmaxf(m, fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP))));
        }
        float s = 0.f, s1 = 0.f, s2 = 0.f;
        for (int k = -krt; k <= krt; ++k) {
            const float th = theta + RP_TWO_PI * float(k);
            const float x = th * th;
            const float e = __expf(fmaf(-a, x, __logf(fmaxf(x, RP_CLAMP))) - m);
            const float dl = x >= RP_CLAMP ? __fdividef(2.f, th) : 0.f;
            s += e;
            s1 = fmaf(e, fmaf(-2.f * a, th, dl), s1);
            s2 = fmaf(e, x, s2);
        }
        const float inv = 1.f / s;
        *d1 = s1 * inv;
        *e2 = s2 * inv;
    }
}

// ------------------------------------------------------------------ staging of broadcast rows
// rows [i0, i0+rows) of a (n,B,W) view of a (B,W) tensor: contiguous unless the tile wraps.
template <int W>
__device__ __forceinline__ void stage_bcast(float* __restrict__ dst, const float* __restrict__ src,
                                            int64_t i0, int rows, int64_t B) {
    // n == 1 (the training case): rows are not broadcast and the emulated 64-bit modulo is skipped
    const int64_t b0 = i0 < B ? i0 : i0 % B;
    if (b0 + rows <= B) {
        tile_g2s(dst, src + b0 * W, rows * W);
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
template <int KT, bool EULER>
__global__ void __launch_bounds__(RP_TILE)
so3_reparam_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps,
                       float* __restrict__ z, float* __restrict__ angles, float* __restrict__ log_q, int64_t total,
                       int64_t B, int krt) {
    __shared__ __align__(16) float s_m[RP_TILE * 9];   // mu in, z out (same row, same thread)
    __shared__ __align__(16) float s_s[RP_TILE * 3];
    __shared__ __align__(16) float s_e[RP_TILE * 3];   // eps in, Euler angles out
    const int64_t i0 = int64_t(blockIdx.x) * RP_TILE;
    const int rows = int(min(int64_t(RP_TILE), total - i0));
    stage_bcast<9>(s_m, mu, i0, rows, B);
    stage_bcast<3>(s_s, sigma, i0, rows, B);
    tile_g2s(s_e, eps + i0 * 3, rows * 3);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        float m[9], sg[3], v[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = s_m[t * 9 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = s_s[t * 3 + j]; v[j] = s_e[t * 3 + j] * sg[j]; }
        RodriguesCtx<float> k;
        rodrigues_ctx(v, k);
        float R[9], zr[9];
        axis_angle_matrix(k.u, k.s, k.w, R);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                zr[r * 3 + c] = fmaf(m[r * 3], R[c], fmaf(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
                s_m[t * 9 + r * 3 + c] = zr[r * 3 + c];
            }
        if (EULER) {
            float q[4], e[3];
            mat_to_quat_fwd(zr, q);
            quat_to_eazyz_fwd(q, e);
#pragma unroll
            for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = e[j];
        }
        if (log_q != nullptr) {
            const float q0 = k.u[0] / sg[0], q1 = k.u[1] / sg[1], q2 = k.u[2] / sg[2];
            const float a = 0.5f * (q0 * q0 + q1 * q1 + q2 * q2);
            const float lse = winding_lse<KT>(k.theta, a, krt);
            const float den = fmaxf(2.f * k.w, RP_CLAMP);
            log_q[i0 + t] = lse - (logf(sg[0]) + logf(sg[1]) + logf(sg[2])) - RP_LOG_2PI_1P5 - logf(den);
        }
    }
    __syncthreads();
    if (z != nullptr) tile_s2g(z + i0 * 9, s_m, rows * 9);
    if (EULER) tile_s2g(angles + i0 * 3, s_e, rows * 3);
}

// ------------------------------------------------------------------ backward
// per-sample gradients: g_mu (total,9), g_sigma (total,3); for n > 1 the caller sums over n
// (lv_sum_leading_f32).  EULER: the upstream gradient arrives (also) as g_angles and is pulled back
// through matrix -> quaternion -> Euler on the recomputed z.
template <int KT, bool EULER>
__global__ void __launch_bounds__(RP_TILE)
so3_reparam_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, const float* __restrict__ eps,
                       const float* __restrict__ gz, const float* __restrict__ gangles, const float* __restrict__ glq,
                       float* __restrict__ gmu, float* __restrict__ gsigma, int64_t total, int64_t B, int krt) {
    __shared__ __align__(16) float s_m[RP_TILE * 9];
    __shared__ __align__(16) float s_g[RP_TILE * 9];   // gz in, g_mu out
    __shared__ __align__(16) float s_s[RP_TILE * 3];
    __shared__ __align__(16) float s_e[RP_TILE * 3];   // eps in, g_sigma out
    __shared__ __align__(16) float s_a[EULER ? RP_TILE * 3 : 4];   // g_angles in
    const int64_t i0 = int64_t(blockIdx.x) * RP_TILE;
    const int rows = int(min(int64_t(RP_TILE), total - i0));
    stage_bcast<9>(s_m, mu, i0, rows, B);
    stage_bcast<3>(s_s, sigma, i0, rows, B);
    tile_g2s(s_e, eps + i0 * 3, rows * 3);
    if (gz != nullptr) tile_g2s(s_g, gz + i0 * 9, rows * 9);
    if (EULER) tile_g2s(s_a, gangles + i0 * 3, rows * 3);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        float m[9], G[9], sg[3], ep[3], v[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) { m[j] = s_m[t * 9 + j]; G[j] = gz != nullptr ? s_g[t * 9 + j] : 0.f; }
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = s_s[t * 3 + j]; ep[j] = s_e[t * 3 + j]; v[j] = ep[j] * sg[j]; }
        RodriguesCtx<float> k;
        rodrigues_ctx(v, k);
        float R[9];
        axis_angle_matrix(k.u, k.s, k.w, R);
        if (EULER) {
            float zr[9], q[4], ge[3], gq[4], gze[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    zr[r * 3 + c] = fmaf(m[r * 3], R[c], fmaf(m[r * 3 + 1], R[3 + c], m[r * 3 + 2] * R[6 + c]));
#pragma unroll
            for (int j = 0; j < 3; ++j) ge[j] = s_a[t * 3 + j];
            mat_to_quat_fwd(zr, q);
            quat_to_eazyz_bwd(q, ge, gq);
            mat_to_quat_bwd(zr, gq, gze);
#pragma unroll
            for (int j = 0; j < 9; ++j) G[j] += gze[j];
        }
        // z = mu R:  g_mu = gz R^T,  g_R = mu^T gz
        float gR[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                s_g[t * 9 + r * 3 + c] = fmaf(G[r * 3], R[c * 3], fmaf(G[r * 3 + 1], R[c * 3 + 1], G[r * 3 + 2] * R[c * 3 + 2]));
                gR[r * 3 + c] = fmaf(m[r], G[c], fmaf(m[3 + r], G[3 + c], m[6 + r] * G[6 + c]));
            }
        float gth, gu[3];
        rodrigues_bwd_theta_u(k, gR, &gth, gu);
        float gs_direct[3] = {0.f, 0.f, 0.f};
        const float gl = glq != nullptr ? glq[i0 + t] : 0.f;
        if (glq != nullptr) {
            const float is0 = 1.f / sg[0], is1 = 1.f / sg[1], is2 = 1.f / sg[2];
            const float q0 = k.u[0] * is0, q1 = k.u[1] * is1, q2 = k.u[2] * is2;
            const float a = 0.5f * (q0 * q0 + q1 * q1 + q2 * q2);
            float d1, e2;
            winding_grad<KT>(k.theta, a, krt, &d1, &e2);
            // - d/dtheta log max(2w, c), 2w = 2 - 2cos: (2 sin)/(2w) = s/w where not clamped
            const float dden = (2.f * k.w >= RP_CLAMP) ? k.s / k.w : 0.f;
            gth = fmaf(gl, d1 - dden, gth);
            // d log_q / d a = -e2 ; d a / d u_i = u_i / sigma_i^2 ; d a / d sigma_i = -u_i^2 / sigma_i^3
            gu[0] = fmaf(-gl * e2, q0 * is0, gu[0]);
            gu[1] = fmaf(-gl * e2, q1 * is1, gu[1]);
            gu[2] = fmaf(-gl * e2, q2 * is2, gu[2]);
            gs_direct[0] = gl * (e2 * q0 * q0 - 1.f) * is0;
            gs_direct[1] = gl * (e2 * q1 * q1 - 1.f) * is1;
            gs_direct[2] = gl * (e2 * q2 * q2 - 1.f) * is2;
        }
        float gv[3];
        theta_u_to_v(k, gth, gu, gv);
#pragma unroll
        for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = fmaf(gv[j], ep[j], gs_direct[j]);
    }
    __syncthreads();
    tile_s2g(gmu + i0 * 9, s_g, rows * 9);
    tile_s2g(gsigma + i0 * 3, s_e, rows * 3);
}

}  // namespace lv

// ====================================================================== C ABI
static int reparam_check(const char* name, int64_t n, int64_t B, int k) {
    if (n < 0 || B < 0 || k < 0) { lv::set_error("%s: negative size", name); return LV_ERR_ARG; }
    if (k > 64) { lv::set_error("%s: k=%d winding terms unsupported (max 64)", name, k); return LV_ERR_UNSUPPORTED; }
    if ((n * B + lv::RP_TILE - 1) / lv::RP_TILE > 0x7fffffffLL) { lv::set_error("%s: too many samples", name); return LV_ERR_ARG; }
    return LV_OK;
}

template <bool EULER>
static int reparam_fwd(const char* name, const float* mu, const float* sigma, const float* eps, float* z, float* angles,
                       float* log_q, int64_t n, int64_t B, int k, void* stream) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || !eps || (EULER ? !angles : !z)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = unsigned((total + lv::RP_TILE - 1) / lv::RP_TILE);
    if (k == 3) lv::so3_reparam_fwd_kernel<3, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    else if (k == 10) lv::so3_reparam_fwd_kernel<10, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    else lv::so3_reparam_fwd_kernel<0, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, z, angles, log_q, total, B, k);
    return lv::check_launch(name);
}

template <bool EULER>
static int reparam_bwd(const char* name, const float* mu, const float* sigma, const float* eps, const float* gz,
                       const float* gangles, const float* glq, float* gmu, float* gsigma, int64_t n, int64_t B, int k,
                       void* stream) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || !eps || !gmu || !gsigma || (EULER && !gangles)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const unsigned grid = unsigned((total + lv::RP_TILE - 1) / lv::RP_TILE);
    if (k == 3) lv::so3_reparam_bwd_kernel<3, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    else if (k == 10) lv::so3_reparam_bwd_kernel<10, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    else lv::so3_reparam_bwd_kernel<0, EULER><<<grid, lv::RP_TILE, 0, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, total, B, k);
    return lv::check_launch(name);
}

extern "C" int lv_so3_reparam_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z, float* log_q,
                                      int64_t n, int64_t B, int k, void* stream) {
    return reparam_fwd<false>("so3_reparam_fwd", mu, sigma, eps, z, nullptr, log_q, n, B, k, stream);
}

extern "C" int lv_so3_reparam_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz,
                                      const float* glq, float* gmu, float* gsigma, int64_t n, int64_t B, int k,
                                      void* stream) {
    return reparam_bwd<false>("so3_reparam_bwd", mu, sigma, eps, gz, nullptr, glq, gmu, gsigma, n, B, k, stream);
}

// reparameterize fused with matrix -> ZYZ Euler (the pose the action decoder consumes); z is optional
extern "C" int lv_so3_reparam_eazyz_fwd_f32(const float* mu, const float* sigma, const float* eps, float* z, float* angles,
                                            float* log_q, int64_t n, int64_t B, int k, void* stream) {
    return reparam_fwd<true>("so3_reparam_eazyz_fwd", mu, sigma, eps, z, angles, log_q, n, B, k, stream);
}

extern "C" int lv_so3_reparam_eazyz_bwd_f32(const float* mu, const float* sigma, const float* eps, const float* gz,
                                            const float* gangles, const float* glq, float* gmu, float* gsigma, int64_t n,
                                            int64_t B, int k, void* stream) {
    return reparam_bwd<true>("so3_reparam_eazyz_bwd", mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, stream);
}
