// Shared device helpers for the lie-vae SO(3) hot-path kernels (sm_100a).
//
// Everything here is per-sample register math: the 3-vector / 3x3 / quaternion
// maps of the reference's lie_tools.py, forward and hand-derived backward,
// templated on the scalar type (float in production, double for the FP64-capable
// elementwise family).  The kernels in elementwise.cu / reparam.cu stage AoS rows
// through shared memory so global accesses stay coalesced and 128-bit wide.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define LV_OK 0
#define LV_ERR_ARG -1
#define LV_ERR_UNSUPPORTED -2
#define LV_ERR_ALIGN -3

namespace lv {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---------------------------------------------------------------- scalar math by type
template <typename T> struct Sc;
template <> struct Sc<float> {
    static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
    static __device__ __forceinline__ float rsqrt(float x) { return rsqrtf(x); }
    static __device__ __forceinline__ void sincos(float x, float* s, float* c) { sincosf(x, s, c); }
    static __device__ __forceinline__ float acos(float x) { return acosf(x); }
    static __device__ __forceinline__ float atan2(float y, float x) { return atan2f(y, x); }
    static __device__ __forceinline__ float tanh(float x) { return tanhf(x); }
    static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
    static __device__ __forceinline__ float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float min(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float rint(float x) { return rintf(x); }
    static __device__ __forceinline__ float log(float x) { return logf(x); }
    static __device__ __forceinline__ float exp(float x) { return expf(x); }
};
template <> struct Sc<double> {
    static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
    static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
    static __device__ __forceinline__ void sincos(double x, double* s, double* c) { ::sincos(x, s, c); }
    static __device__ __forceinline__ double acos(double x) { return ::acos(x); }
    static __device__ __forceinline__ double atan2(double y, double x) { return ::atan2(y, x); }
    static __device__ __forceinline__ double tanh(double x) { return ::tanh(x); }
    static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
    static __device__ __forceinline__ double fma(double a, double b, double c) { return ::fma(a, b, c); }
    static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
    static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
    static __device__ __forceinline__ double rint(double x) { return ::rint(x); }
    static __device__ __forceinline__ double log(double x) { return ::log(x); }
    static __device__ __forceinline__ double exp(double x) { return ::exp(x); }
};

// ---------------------------------------------------------------- coalesced AoS staging
// A CTA owns `rows` consecutive rows of W scalars; the tile is one contiguous
// span of global memory, copied with 16-byte accesses when the span start is
// 16-byte aligned and scalar (still fully coalesced) accesses otherwise.
// Asynchronous global->shared tile copy (cp.async / LDGSTS): every thread puts ALL of its pieces in
// flight before anyone waits, so a CTA has its whole tile outstanding at once (a load->store loop
// through registers keeps only a few 16-byte requests per thread in flight and is latency-bound).
// 16-byte pieces bypass L1 (.cg: each byte is used once); spans that are not 16-byte aligned fall
// back to 4-byte pieces (.ca), still asynchronous and fully coalesced.
// Complete with tile_async_wait() followed by __syncthreads().
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void tile_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <typename T>
__device__ __forceinline__ void tile_g2s(T* __restrict__ dst, const T* __restrict__ src, int count) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte scalars");
    constexpr int V = 16 / sizeof(T);
    const int tid = threadIdx.x, nt = blockDim.x;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        const int nv = count / V;
        for (int i = tid; i < nv; i += nt) cp_async16(dst + i * V, src + i * V);
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + nv * V);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst + nv * V);
        const int rest = (count - nv * V) * int(sizeof(T) / 4);
        for (int i = tid; i < rest; i += nt) cp_async4(d32 + i, s32 + i);
    } else {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
        const int n32 = count * int(sizeof(T) / 4);
        for (int i = tid; i < n32; i += nt) cp_async4(d32 + i, s32 + i);
    }
}

template <typename T>
__device__ __forceinline__ void tile_s2g(T* __restrict__ dst, const T* __restrict__ src, int count) {
    constexpr int V = 16 / sizeof(T);
    const int tid = threadIdx.x, nt = blockDim.x;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int nv = count / V;
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        for (int i = tid; i < nv; i += nt) d4[i] = s4[i];
        for (int i = nv * V + tid; i < count; i += nt) dst[i] = src[i];
    } else {
        for (int i = tid; i < count; i += nt) dst[i] = src[i];
    }
}

// Full-tile variants: element count and thread count are compile-time constants, so the copy loops unroll completely,
// every piece's offset is an immediate and no 64-bit index arithmetic is left (the run-time-count versions above spend
// ~40 % of a latent kernel's instructions on loop control and address math).  Same alignment rule as above.
template <typename T, int COUNT, int NT>
__device__ __forceinline__ void tile_g2s_full(T* __restrict__ dst, const T* __restrict__ src) {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "4- or 8-byte scalars");
    constexpr int V = 16 / sizeof(T), NV = COUNT / V, REST32 = (COUNT - NV * V) * int(sizeof(T) / 4);
    constexpr int N32 = COUNT * int(sizeof(T) / 4);
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
#pragma unroll
        for (int k = 0; k < (NV + NT - 1) / NT; ++k) {
            const int i = tid + k * NT;
            if ((k + 1) * NT <= NV || i < NV) cp_async16(dst + i * V, src + i * V);
        }
        if (REST32 > 0 && tid < REST32)
            cp_async4(reinterpret_cast<uint32_t*>(dst + NV * V) + tid, reinterpret_cast<const uint32_t*>(src + NV * V) + tid);
    } else {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
        uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
#pragma unroll
        for (int k = 0; k < (N32 + NT - 1) / NT; ++k) {
            const int i = tid + k * NT;
            if ((k + 1) * NT <= N32 || i < N32) cp_async4(d32 + i, s32 + i);
        }
    }
}
template <typename T, int COUNT, int NT>
__device__ __forceinline__ void tile_s2g_full(T* __restrict__ dst, const T* __restrict__ src) {
    constexpr int V = 16 / sizeof(T), NV = COUNT / V, REST = COUNT - NV * V;
    const int tid = threadIdx.x;
    if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
#pragma unroll
        for (int k = 0; k < (NV + NT - 1) / NT; ++k) {
            const int i = tid + k * NT;
            if ((k + 1) * NT <= NV || i < NV) d4[i] = s4[i];
        }
        if (REST > 0 && tid < REST) dst[NV * V + tid] = src[NV * V + tid];
    } else {
#pragma unroll
        for (int k = 0; k < (COUNT + NT - 1) / NT; ++k) {
            const int i = tid + k * NT;
            if ((k + 1) * NT <= COUNT || i < COUNT) dst[i] = src[i];
        }
    }
}

// ---------------------------------------------------------------- mbarrier / TMA bulk-copy helpers (sm_100a)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the warp sleeps in the barrier unit instead of spinning through the issue slots its
// CTA's other warps need (ncu, first degree-specialised Wigner backward: 12 % of all issued instructions were wait loops)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory");
}
// global -> shared, contiguous span (16-byte aligned on both sides, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void tma_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global, same constraints; part of the thread's current bulk group (tma_store_commit_wait)
__device__ __forceinline__ void tma_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// commit the bulk stores issued so far and wait until the copy engine has READ their shared-memory sources
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until all but the N most recent bulk groups of this thread have had their shared-memory sources read
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---------------------------------------------------------------- 3x3 helpers (row-major r[9])
// hat(u): [[0,-u2,u1],[u2,0,-u0],[-u1,u0,0]]      (lie_tools.py:17-43)
// <G, hat(w)> = w . axial(G),  axial(G) = (G21-G12, G02-G20, G10-G01)
template <typename T>
__device__ __forceinline__ void axial(const T* G, T* a) {
    a[0] = G[7] - G[5];
    a[1] = G[2] - G[6];
    a[2] = G[3] - G[1];
}

// R = I + s*hat(u) + w*(u u^T - (u.u) I)   -- the body shared by rodrigues / s2s1rodrigues
// (lie_tools.py:62-63, :75-76).  For a unit axis this is the rotation by the angle with
// sin = s, 1-cos = w.
template <typename T>
__device__ __forceinline__ void axis_angle_matrix(const T* u, T s, T w, T* R) {
    const T uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    const T d = T(1) - w * uu;
    const T w01 = w * u[0] * u[1], w02 = w * u[0] * u[2], w12 = w * u[1] * u[2];
    R[0] = d + w * u[0] * u[0];  R[1] = w01 - s * u[2];       R[2] = w02 + s * u[1];
    R[3] = w01 + s * u[2];       R[4] = d + w * u[1] * u[1];  R[5] = w12 - s * u[0];
    R[6] = w02 - s * u[1];       R[7] = w12 + s * u[0];       R[8] = d + w * u[2] * u[2];
}

// Backward of axis_angle_matrix w.r.t. (u, s, w) for upstream G (3x3):
//   gs = u . axial(G);  gw = u^T G u - (u.u) tr G;  gu = s*axial(G) + w*((G+G^T)u - 2 tr(G) u)
template <typename T>
__device__ __forceinline__ void axis_angle_matrix_bwd(const T* u, T s, T w, const T* G, T* gu, T* gs, T* gw) {
    T a[3];
    axial(G, a);
    const T tr = G[0] + G[4] + G[8];
    const T uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    T Su[3];   // (G + G^T) u
    Su[0] = T(2) * G[0] * u[0] + (G[1] + G[3]) * u[1] + (G[2] + G[6]) * u[2];
    Su[1] = (G[1] + G[3]) * u[0] + T(2) * G[4] * u[1] + (G[5] + G[7]) * u[2];
    Su[2] = (G[2] + G[6]) * u[0] + (G[5] + G[7]) * u[1] + T(2) * G[8] * u[2];
    *gs = u[0] * a[0] + u[1] * a[1] + u[2] * a[2];
    *gw = T(0.5) * (u[0] * Su[0] + u[1] * Su[1] + u[2] * Su[2]) - uu * tr;
#pragma unroll
    for (int i = 0; i < 3; ++i) gu[i] = s * a[i] + w * (Su[i] - T(2) * tr * u[i]);
}

// ---------------------------------------------------------------- rodrigues (lie_tools.py:56-64)
// theta = |v|, u = v/theta, R = I + sin(theta) hat(u) + (1-cos(theta)) hat(u)^2.
// sin/cos come from the half angle so that 1-cos = 2 sin^2(theta/2) has no cancellation.
// At v = 0 the reference returns NaN (0/0); here u := 0 and R = I (documented improvement).
template <typename T>
struct RodriguesCtx {
    T u[3], theta, inv_theta, s, c, w;   // s = sin, c = cos, w = 1 - cos
};

template <typename T>
__device__ __forceinline__ void rodrigues_ctx(const T* v, RodriguesCtx<T>& k) {
    const T t2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    k.theta = Sc<T>::sqrt(t2);
    k.inv_theta = k.theta > T(0) ? T(1) / k.theta : T(0);
    k.u[0] = v[0] * k.inv_theta; k.u[1] = v[1] * k.inv_theta; k.u[2] = v[2] * k.inv_theta;
    T sh, ch;
    Sc<T>::sincos(T(0.5) * k.theta, &sh, &ch);
    k.s = T(2) * sh * ch;
    k.w = T(2) * sh * sh;
    k.c = T(1) - k.w;
}

template <typename T>
__device__ __forceinline__ void rodrigues_fwd(const T* v, T* R) {
    RodriguesCtx<T> k;
    rodrigues_ctx(v, k);
    axis_angle_matrix(k.u, k.s, k.w, R);
}

// gv from upstream G = dL/dR.  g_theta / g_u are also returned for callers that add
// further theta/u dependent terms (the wrapped log-density) before projecting to v.
template <typename T>
__device__ __forceinline__ void rodrigues_bwd_theta_u(const RodriguesCtx<T>& k, const T* G, T* gtheta, T* gu) {
    T gs, gw;
    axis_angle_matrix_bwd(k.u, k.s, k.w, G, gu, &gs, &gw);
    *gtheta = gs * k.c + gw * k.s;      // d sin = cos, d(1-cos) = sin
}

// v = theta*u-parametrisation back to v:  dtheta = u.dv,  du = (dv - u (u.dv)) / theta
template <typename T>
__device__ __forceinline__ void theta_u_to_v(const RodriguesCtx<T>& k, T gtheta, const T* gu, T* gv) {
    const T ug = k.u[0] * gu[0] + k.u[1] * gu[1] + k.u[2] * gu[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) gv[i] = gtheta * k.u[i] + (gu[i] - ug * k.u[i]) * k.inv_theta;
}

template <typename T>
__device__ __forceinline__ void rodrigues_bwd(const T* v, const T* G, T* gv) {
    RodriguesCtx<T> k;
    rodrigues_ctx(v, k);
    T gt, gu[3];
    rodrigues_bwd_theta_u(k, G, &gt, gu);
    theta_u_to_v(k, gt, gu, gv);
}

// ---------------------------------------------------------------- quaternion -> matrix (lie_tools.py:183-192)
template <typename T>
__device__ __forceinline__ void quat_to_mat_fwd(const T* q, T* R) {
    const T inv = Sc<T>::rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const T a = q[0] * inv, b = q[1] * inv, c = q[2] * inv, d = q[3] * inv;
    R[0] = a * a - b * b - c * c + d * d; R[1] = T(2) * (a * b + c * d);         R[2] = T(2) * (a * c - b * d);
    R[3] = T(2) * (a * b - c * d);        R[4] = -a * a + b * b - c * c + d * d; R[5] = T(2) * (b * c + a * d);
    R[6] = T(2) * (a * c + b * d);        R[7] = T(2) * (b * c - a * d);         R[8] = -a * a - b * b + c * c + d * d;
}

template <typename T>
__device__ __forceinline__ void quat_to_mat_bwd(const T* q, const T* G, T* gq) {
    const T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const T inv = Sc<T>::rsqrt(n2);
    const T a = q[0] * inv, b = q[1] * inv, c = q[2] * inv, d = q[3] * inv;
    // gradient w.r.t. the normalised quaternion
    T gn[4];
    gn[0] = T(2) * (a * (G[0] - G[4] - G[8]) + b * (G[1] + G[3]) + c * (G[2] + G[6]) + d * (G[5] - G[7]));
    gn[1] = T(2) * (b * (-G[0] + G[4] - G[8]) + a * (G[1] + G[3]) + c * (G[5] + G[7]) + d * (G[6] - G[2]));
    gn[2] = T(2) * (c * (-G[0] - G[4] + G[8]) + a * (G[2] + G[6]) + b * (G[5] + G[7]) + d * (G[1] - G[3]));
    gn[3] = T(2) * (d * (G[0] + G[4] + G[8]) + a * (G[5] - G[7]) + b * (G[6] - G[2]) + c * (G[1] - G[3]));
    // through q / |q|
    const T dot = a * gn[0] + b * gn[1] + c * gn[2] + d * gn[3];
    gq[0] = (gn[0] - a * dot) * inv; gq[1] = (gn[1] - b * dot) * inv;
    gq[2] = (gn[2] - c * dot) * inv; gq[3] = (gn[3] - d * dot) * inv;
}

// ---------------------------------------------------------------- matrix -> quaternion (lie_tools.py:112-157)
// Four Shepperd candidates; the one with the largest denominator is selected (first
// index on ties, like torch.argmax).  Only the selected branch is evaluated.
template <typename T>
struct ShepperdSel {
    int j;       // selected branch
    T den;       // 0.5*sqrt(1e-6 + |pre_j|)
    T sgn;       // sign(pre_j) (0 at 0, like torch.abs backward)
};

template <typename T>
__device__ __forceinline__ void shepperd_select(const T* r, ShepperdSel<T>& s) {
    const T d0 = r[0], d1 = r[4], d2 = r[8];
    T pre[4] = {T(1) + d0 - d1 - d2, T(1) - d0 + d1 - d2, T(1) - d0 - d1 + d2, T(1) + d0 + d1 + d2};
    // argmax over the rounded denominators themselves, exactly as the reference does
    T den[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) den[i] = T(0.5) * Sc<T>::sqrt(T(1e-6) + Sc<T>::abs(pre[i]));
    int j = 0;
    T best = den[0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (den[i] > best) { best = den[i]; j = i; }
    s.j = j;
    s.den = best;
    const T p = j == 0 ? pre[0] : j == 1 ? pre[1] : j == 2 ? pre[2] : pre[3];
    s.sgn = p > T(0) ? T(1) : (p < T(0) ? T(-1) : T(0));
}

// numerators of the three off-diagonal quaternion components for branch j:
//   j=0: (s01, s02, a12) -> q1,q2,q3     j=1: (s01, s12, a20) -> q0,q2,q3
//   j=2: (s02, s12, a01) -> q0,q1,q3     j=3: (a12, a20, a01) -> q0,q1,q2
template <typename T>
__device__ __forceinline__ void mat_to_quat_fwd(const T* r, T* q) {
    ShepperdSel<T> s;
    shepperd_select(r, s);
    const T f = T(1) / (T(4) * s.den);
    const T s01 = r[1] + r[3], s02 = r[2] + r[6], s12 = r[5] + r[7];
    const T a12 = r[5] - r[7], a20 = r[6] - r[2], a01 = r[1] - r[3];
    if (s.j == 0)      { q[0] = s.den;  q[1] = s01 * f; q[2] = s02 * f; q[3] = a12 * f; }
    else if (s.j == 1) { q[0] = s01 * f; q[1] = s.den;  q[2] = s12 * f; q[3] = a20 * f; }
    else if (s.j == 2) { q[0] = s02 * f; q[1] = s12 * f; q[2] = s.den;  q[3] = a01 * f; }
    else               { q[0] = a12 * f; q[1] = a20 * f; q[2] = a01 * f; q[3] = s.den; }
}

template <typename T>
__device__ __forceinline__ void mat_to_quat_bwd(const T* r, const T* gq, T* gr) {
    ShepperdSel<T> s;
    shepperd_select(r, s);
    const T f = T(1) / (T(4) * s.den);
    const T s01 = r[1] + r[3], s02 = r[2] + r[6], s12 = r[5] + r[7];
    const T a12 = r[5] - r[7], a20 = r[6] - r[2], a01 = r[1] - r[3];
    // q_j = den ; q_other = num * f, f = 1/(4 den)  ->  d q_other / d den = -num f / den
    T gs01 = 0, gs02 = 0, gs12 = 0, ga12 = 0, ga20 = 0, ga01 = 0, gden;
    if (s.j == 0) {
        gs01 = gq[1] * f; gs02 = gq[2] * f; ga12 = gq[3] * f;
        gden = gq[0] - (gq[1] * s01 + gq[2] * s02 + gq[3] * a12) * f / s.den;
    } else if (s.j == 1) {
        gs01 = gq[0] * f; gs12 = gq[2] * f; ga20 = gq[3] * f;
        gden = gq[1] - (gq[0] * s01 + gq[2] * s12 + gq[3] * a20) * f / s.den;
    } else if (s.j == 2) {
        gs02 = gq[0] * f; gs12 = gq[1] * f; ga01 = gq[3] * f;
        gden = gq[2] - (gq[0] * s02 + gq[1] * s12 + gq[3] * a01) * f / s.den;
    } else {
        ga12 = gq[0] * f; ga20 = gq[1] * f; ga01 = gq[2] * f;
        gden = gq[3] - (gq[0] * a12 + gq[1] * a20 + gq[2] * a01) * f / s.den;
    }
    // den = 0.5 sqrt(1e-6 + |pre|)  ->  d den / d pre = sgn / (8 den)
    const T gpre = gden * s.sgn / (T(8) * s.den);
    // pre_j = 1 + e0 r00 + e1 r11 + e2 r22
    const T e0 = (s.j == 0 || s.j == 3) ? T(1) : T(-1);
    const T e1 = (s.j == 1 || s.j == 3) ? T(1) : T(-1);
    const T e2 = (s.j == 2 || s.j == 3) ? T(1) : T(-1);
    gr[0] = e0 * gpre; gr[4] = e1 * gpre; gr[8] = e2 * gpre;
    gr[1] = gs01 + ga01; gr[3] = gs01 - ga01;
    gr[2] = gs02 - ga20; gr[6] = gs02 + ga20;
    gr[5] = gs12 + ga12; gr[7] = gs12 - ga12;
}

// ---------------------------------------------------------------- quaternion -> ZYZ Euler (lie_tools.py:160-175)
template <typename T>
__device__ __forceinline__ void quat_to_eazyz_fwd(const T* q, T* e) {
    const T lo = T(-1.0 + 1e-6), hi = T(1.0 - 1e-6);
    e[0] = Sc<T>::atan2(q[1] * q[2] - q[0] * q[3], q[0] * q[2] + q[1] * q[3]);
    T w = q[3] * q[3] - q[0] * q[0] - q[1] * q[1] + q[2] * q[2];
    w = w < lo ? lo : (w > hi ? hi : w);
    e[1] = Sc<T>::acos(w);
    e[2] = Sc<T>::atan2(q[0] * q[3] + q[1] * q[2], q[1] * q[3] - q[0] * q[2]);
}

template <typename T>
__device__ __forceinline__ void quat_to_eazyz_bwd(const T* q, const T* ge, T* gq) {
    const T lo = T(-1.0 + 1e-6), hi = T(1.0 - 1e-6);
    // alpha = atan2(ya, xa)
    const T ya = q[1] * q[2] - q[0] * q[3], xa = q[0] * q[2] + q[1] * q[3];
    // torch's atan2 backward zeroes the reciprocal at the origin (x = y = 0)
    const T da = xa * xa + ya * ya;
    const T ra = da > T(0) ? ge[0] / da : T(0);
    const T gya = xa * ra, gxa = -ya * ra;
    // gamma = atan2(yc, xc)
    const T yc = q[0] * q[3] + q[1] * q[2], xc = q[1] * q[3] - q[0] * q[2];
    const T dc = xc * xc + yc * yc;
    const T rc = dc > T(0) ? ge[2] / dc : T(0);
    const T gyc = xc * rc, gxc = -yc * rc;
    // beta = acos(clamp(w)); clamp passes gradient on the closed interval
    const T w = q[3] * q[3] - q[0] * q[0] - q[1] * q[1] + q[2] * q[2];
    const T gw = (w >= lo && w <= hi) ? -ge[1] * Sc<T>::rsqrt(T(1) - w * w) : T(0);
    gq[0] = -gya * q[3] + gxa * q[2] + gyc * q[3] - gxc * q[2] - T(2) * gw * q[0];
    gq[1] = gya * q[2] + gxa * q[3] + gyc * q[2] + gxc * q[3] - T(2) * gw * q[1];
    gq[2] = gya * q[1] + gxa * q[0] + gyc * q[1] - gxc * q[0] + T(2) * gw * q[2];
    gq[3] = -gya * q[0] + gxa * q[1] + gyc * q[0] + gxc * q[1] + T(2) * gw * q[3];
}

// ---------------------------------------------------------------- s2s2_gram_schmidt (lie_tools.py:81-89)
template <typename T>
__device__ __forceinline__ void normalize_clamped_bwd(const T* e, T nrm, T cl, const T* ge, T* gx) {
    // x -> x / max(|x|, 1e-5): torch.clamp passes the norm's gradient where |x| >= 1e-5
    if (nrm >= T(1e-5)) {
        const T d = e[0] * ge[0] + e[1] * ge[1] + e[2] * ge[2];
#pragma unroll
        for (int i = 0; i < 3; ++i) gx[i] = (ge[i] - e[i] * d) / cl;
    } else {
#pragma unroll
        for (int i = 0; i < 3; ++i) gx[i] = ge[i] / cl;
    }
}
template <typename T>
__device__ __forceinline__ void s2s2_fwd(const T* v1, const T* v2, T* R) {
        const T c1 = Sc<T>::max(Sc<T>::sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]), T(1e-5));
        T e1[3] = {v1[0] / c1, v1[1] / c1, v1[2] / c1};
        const T p = e1[0] * v2[0] + e1[1] * v2[1] + e1[2] * v2[2];
        T u2[3] = {v2[0] - p * e1[0], v2[1] - p * e1[1], v2[2] - p * e1[2]};
        const T c2 = Sc<T>::max(Sc<T>::sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]), T(1e-5));
        T e2[3] = {u2[0] / c2, u2[1] / c2, u2[2] / c2};
        R[0] = e1[0]; R[1] = e1[1]; R[2] = e1[2];
        R[3] = e2[0]; R[4] = e2[1]; R[5] = e2[2];
        R[6] = e1[1] * e2[2] - e1[2] * e2[1];
        R[7] = e1[2] * e2[0] - e1[0] * e2[2];
        R[8] = e1[0] * e2[1] - e1[1] * e2[0];
}
template <typename T>
__device__ __forceinline__ void s2s2_bwd(const T* v1, const T* v2, const T* G, T* gv1, T* gv2) {
        const T n1 = Sc<T>::sqrt(v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2]);
        const T c1 = Sc<T>::max(n1, T(1e-5));
        T e1[3] = {v1[0] / c1, v1[1] / c1, v1[2] / c1};
        const T p = e1[0] * v2[0] + e1[1] * v2[1] + e1[2] * v2[2];
        T u2[3] = {v2[0] - p * e1[0], v2[1] - p * e1[1], v2[2] - p * e1[2]};
        const T n2 = Sc<T>::sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
        const T c2 = Sc<T>::max(n2, T(1e-5));
        T e2[3] = {u2[0] / c2, u2[1] / c2, u2[2] / c2};
        const T* g1 = G; const T* g2 = G + 3; const T* g3 = G + 6;
        // e3 = e1 x e2
        T ge1[3] = {g1[0] + (e2[1] * g3[2] - e2[2] * g3[1]), g1[1] + (e2[2] * g3[0] - e2[0] * g3[2]),
                    g1[2] + (e2[0] * g3[1] - e2[1] * g3[0])};
        T ge2[3] = {g2[0] + (g3[1] * e1[2] - g3[2] * e1[1]), g2[1] + (g3[2] * e1[0] - g3[0] * e1[2]),
                    g2[2] + (g3[0] * e1[1] - g3[1] * e1[0])};
        T gu2[3];
        normalize_clamped_bwd(e2, n2, c2, ge2, gu2);
        // u2 = v2 - p e1, p = e1.v2
        const T gp = -(gu2[0] * e1[0] + gu2[1] * e1[1] + gu2[2] * e1[2]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gv2[i] = gu2[i] + gp * e1[i];
            ge1[i] += -p * gu2[i] + gp * v2[i];
        }
        normalize_clamped_bwd(e1, n1, c1, ge1, gv1);
}

}  // namespace lv
