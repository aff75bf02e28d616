// Elementwise SO(3) maps (forward + backward), float and double.
//
// Replaces, one fused kernel each, the ATen chains behind the reference's
// lie_tools.py functions.  All of them are HBM-bound:
// one thread per sample, register math from common.cuh, AoS rows (3/4/9 scalars)
// staged through shared memory so that every global access is a contiguous,
// 128-bit-wide span per CTA.  Row strides 3 and 9 are odd -> conflict-free LDS;
// strides 2 and 4 use 64/128-bit LDS.
#include "common.cuh"

namespace lv {

// ------------------------------------------------------------------ row <-> smem helpers
template <typename T, int W>
__device__ __forceinline__ void row_load(const T* __restrict__ s, T* r) {
    if constexpr ((W * sizeof(T)) % 16 == 0) {
        constexpr int NV = W * sizeof(T) / 16;
        const int4* s4 = reinterpret_cast<const int4*>(s);
        int4* r4 = reinterpret_cast<int4*>(r);
#pragma unroll
        for (int i = 0; i < NV; ++i) r4[i] = s4[i];
    } else if constexpr ((W * sizeof(T)) % 8 == 0) {
        constexpr int NV = W * sizeof(T) / 8;
        const int2* s2 = reinterpret_cast<const int2*>(s);
        int2* r2 = reinterpret_cast<int2*>(r);
#pragma unroll
        for (int i = 0; i < NV; ++i) r2[i] = s2[i];
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) r[i] = s[i];
    }
}
template <typename T, int W>
__device__ __forceinline__ void row_store(T* __restrict__ s, const T* r) {
    if constexpr ((W * sizeof(T)) % 16 == 0) {
        constexpr int NV = W * sizeof(T) / 16;
        int4* s4 = reinterpret_cast<int4*>(s);
        const int4* r4 = reinterpret_cast<const int4*>(r);
#pragma unroll
        for (int i = 0; i < NV; ++i) s4[i] = r4[i];
    } else if constexpr ((W * sizeof(T)) % 8 == 0) {
        constexpr int NV = W * sizeof(T) / 8;
        int2* s2 = reinterpret_cast<int2*>(s);
        const int2* r2 = reinterpret_cast<const int2*>(r);
#pragma unroll
        for (int i = 0; i < NV; ++i) s2[i] = r2[i];
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) s[i] = r[i];
    }
}

__host__ __device__ constexpr int align16(int bytes) { return (bytes + 15) & ~15; }

// ------------------------------------------------------------------ the generic row kernel
// Op: static constexpr int I0,I1,I2 (input row widths, 0 = unused), O0,O1 (outputs);
//     static __device__ void run(const T* i0, const T* i1, const T* i2, T* o0, T* o1)
template <typename T, typename Op, int TILE>
__global__ void __launch_bounds__(TILE)
row_kernel(const T* __restrict__ g0, const T* __restrict__ g1, const T* __restrict__ g2,
           T* __restrict__ h0, T* __restrict__ h1, int64_t n) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int B0 = align16(TILE * Op::I0 * sizeof(T));
    constexpr int B1 = align16(TILE * Op::I1 * sizeof(T));
    constexpr int B2 = align16(TILE * Op::I2 * sizeof(T));
    constexpr int C0 = align16(TILE * Op::O0 * sizeof(T));
    T* s0 = reinterpret_cast<T*>(smem_raw);
    T* s1 = reinterpret_cast<T*>(smem_raw + B0);
    T* s2 = reinterpret_cast<T*>(smem_raw + B0 + B1);
    T* t0 = reinterpret_cast<T*>(smem_raw + B0 + B1 + B2);
    T* t1 = reinterpret_cast<T*>(smem_raw + B0 + B1 + B2 + C0);

    const int64_t row0 = int64_t(blockIdx.x) * TILE;
    const int rows = int(min(int64_t(TILE), n - row0));
    const bool full = rows == TILE;       // every CTA but the last: compile-time copy loops
    if (full) {
        if constexpr (Op::I0 > 0) tile_g2s_full<T, TILE * Op::I0, TILE>(s0, g0 + row0 * Op::I0);
        if constexpr (Op::I1 > 0) tile_g2s_full<T, TILE * Op::I1, TILE>(s1, g1 + row0 * Op::I1);
        if constexpr (Op::I2 > 0) tile_g2s_full<T, TILE * Op::I2, TILE>(s2, g2 + row0 * Op::I2);
    } else {
        if constexpr (Op::I0 > 0) tile_g2s(s0, g0 + row0 * Op::I0, rows * Op::I0);
        if constexpr (Op::I1 > 0) tile_g2s(s1, g1 + row0 * Op::I1, rows * Op::I1);
        if constexpr (Op::I2 > 0) tile_g2s(s2, g2 + row0 * Op::I2, rows * Op::I2);
    }
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        alignas(16) T a0[Op::I0 > 0 ? Op::I0 : 1];
        alignas(16) T a1[Op::I1 > 0 ? Op::I1 : 1];
        alignas(16) T a2[Op::I2 > 0 ? Op::I2 : 1];
        alignas(16) T o0[Op::O0 > 0 ? Op::O0 : 1];
        alignas(16) T o1[Op::O1 > 0 ? Op::O1 : 1];
        if constexpr (Op::I0 > 0) row_load<T, Op::I0>(s0 + t * Op::I0, a0);
        if constexpr (Op::I1 > 0) row_load<T, Op::I1>(s1 + t * Op::I1, a1);
        if constexpr (Op::I2 > 0) row_load<T, Op::I2>(s2 + t * Op::I2, a2);
        Op::run(a0, a1, a2, o0, o1);
        if constexpr (Op::O0 > 0) row_store<T, Op::O0>(t0 + t * Op::O0, o0);
        if constexpr (Op::O1 > 0) row_store<T, Op::O1>(t1 + t * Op::O1, o1);
    }
    __syncthreads();
    if (full) {
        if constexpr (Op::O0 > 0) tile_s2g_full<T, TILE * Op::O0, TILE>(h0 + row0 * Op::O0, t0);
        if constexpr (Op::O1 > 0) tile_s2g_full<T, TILE * Op::O1, TILE>(h1 + row0 * Op::O1, t1);
    } else {
        if constexpr (Op::O0 > 0) tile_s2g(h0 + row0 * Op::O0, t0, rows * Op::O0);
        if constexpr (Op::O1 > 0) tile_s2g(h1 + row0 * Op::O1, t1, rows * Op::O1);
    }
}

template <typename T, typename Op>
int launch_rows(const T* g0, const T* g1, const T* g2, T* h0, T* h1, int64_t n, cudaStream_t st, const char* name) {
    if (n < 0) { set_error("%s: negative row count", name); return LV_ERR_ARG; }
    if (n == 0) return LV_OK;
    if ((Op::I0 > 0 && !g0) || (Op::I1 > 0 && !g1) || (Op::I2 > 0 && !g2) || (Op::O0 > 0 && !h0) || (Op::O1 > 0 && !h1)) {
        set_error("%s: null pointer", name);
        return LV_ERR_ARG;
    }
    constexpr int TILE = sizeof(T) == 4 ? 256 : 128;
    constexpr int SMEM = align16(TILE * Op::I0 * sizeof(T)) + align16(TILE * Op::I1 * sizeof(T)) +
                         align16(TILE * Op::I2 * sizeof(T)) + align16(TILE * Op::O0 * sizeof(T)) +
                         align16(TILE * Op::O1 * sizeof(T));
    static_assert(SMEM <= 48 * 1024, "row kernel tile exceeds the static shared-memory window");
    const int64_t blocks = (n + TILE - 1) / TILE;
    if (blocks > 0x7fffffffLL) { set_error("%s: too many rows", name); return LV_ERR_ARG; }
    row_kernel<T, Op, TILE><<<unsigned(blocks), TILE, SMEM, st>>>(g0, g1, g2, h0, h1, n);
    return check_launch(name);
}

// ------------------------------------------------------------------ ops
template <typename T> struct HatFwd {   // lie_tools.py:17-43
    static constexpr int I0 = 3, I1 = 0, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* v, const T*, const T*, T* X, T*) {
        X[0] = 0; X[1] = -v[2]; X[2] = v[1]; X[3] = v[2]; X[4] = 0; X[5] = -v[0]; X[6] = -v[1]; X[7] = v[0]; X[8] = 0;
    }
};
template <typename T> struct HatBwd {
    static constexpr int I0 = 9, I1 = 0, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* g, const T*, const T*, T* gv, T*) { axial(g, gv); }
};
template <typename T> struct VeeFwd {   // lie_tools.py:46-53
    static constexpr int I0 = 9, I1 = 0, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* X, const T*, const T*, T* v, T*) {
        v[0] = -X[5]; v[1] = X[2]; v[2] = -X[1];
    }
};
template <typename T> struct VeeBwd {
    static constexpr int I0 = 3, I1 = 0, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* g, const T*, const T*, T* gX, T*) {
#pragma unroll
        for (int i = 0; i < 9; ++i) gX[i] = 0;
        gX[5] = -g[0]; gX[2] = g[1]; gX[1] = -g[2];
    }
};
template <typename T> struct RodriguesFwd {   // lie_tools.py:56-64
    static constexpr int I0 = 3, I1 = 0, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* v, const T*, const T*, T* R, T*) { rodrigues_fwd(v, R); }
};
template <typename T> struct RodriguesBwd {
    static constexpr int I0 = 3, I1 = 9, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* v, const T* G, const T*, T* gv, T*) { rodrigues_bwd(v, G, gv); }
};
template <typename T> struct LogMapFwd {   // lie_tools.py:100-109, batched
    static constexpr int I0 = 9, I1 = 0, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T*, const T*, T* X, T*) {
        const T c = T(0.5) * (R[0] + R[4] + R[8] - T(1));
        const T theta = Sc<T>::acos(c);
        T s, cc;
        Sc<T>::sincos(theta, &s, &cc);
        const T f = T(0.5) * theta / s;
        X[0] = 0; X[4] = 0; X[8] = 0;
        X[1] = f * (R[1] - R[3]); X[3] = -X[1];
        X[2] = f * (R[2] - R[6]); X[6] = -X[2];
        X[5] = f * (R[5] - R[7]); X[7] = -X[5];
    }
};
template <typename T> struct LogMapBwd {
    static constexpr int I0 = 9, I1 = 9, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T* G, const T*, T* gR, T*) {
        const T c = T(0.5) * (R[0] + R[4] + R[8] - T(1));
        const T theta = Sc<T>::acos(c);
        T s, cc;
        Sc<T>::sincos(theta, &s, &cc);
        const T f = theta / s;
        // dL/df = <G, (R - R^T)/2>
        const T gf = T(0.5) * ((G[1] - G[3]) * (R[1] - R[3]) + (G[2] - G[6]) * (R[2] - R[6]) + (G[5] - G[7]) * (R[5] - R[7]));
        // df/dtheta = (sin - theta cos)/sin^2 ; dtheta/dtr = -1/(2 sqrt(1-c^2))
        const T gtr = gf * (s - theta * cc) / (s * s) * (T(-0.5) * Sc<T>::rsqrt(T(1) - c * c));
        const T h = T(0.5) * f;
        gR[0] = gtr; gR[4] = gtr; gR[8] = gtr;
        gR[1] = h * (G[1] - G[3]); gR[3] = -gR[1];
        gR[2] = h * (G[2] - G[6]); gR[6] = -gR[2];
        gR[5] = h * (G[5] - G[7]); gR[7] = -gR[5];
    }
};
template <typename T> struct QuatToMatFwd {   // lie_tools.py:183-192
    static constexpr int I0 = 4, I1 = 0, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* q, const T*, const T*, T* R, T*) { quat_to_mat_fwd(q, R); }
};
template <typename T> struct QuatToMatBwd {
    static constexpr int I0 = 4, I1 = 9, I2 = 0, O0 = 4, O1 = 0;
    static __device__ __forceinline__ void run(const T* q, const T* G, const T*, T* gq, T*) { quat_to_mat_bwd(q, G, gq); }
};
template <typename T> struct MatToQuatFwd {   // lie_tools.py:112-157
    static constexpr int I0 = 9, I1 = 0, I2 = 0, O0 = 4, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T*, const T*, T* q, T*) { mat_to_quat_fwd(R, q); }
};
template <typename T> struct MatToQuatBwd {
    static constexpr int I0 = 9, I1 = 4, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T* gq, const T*, T* gR, T*) { mat_to_quat_bwd(R, gq, gR); }
};
template <typename T> struct QuatToEazyzFwd {   // lie_tools.py:160-175
    static constexpr int I0 = 4, I1 = 0, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* q, const T*, const T*, T* e, T*) { quat_to_eazyz_fwd(q, e); }
};
template <typename T> struct QuatToEazyzBwd {
    static constexpr int I0 = 4, I1 = 3, I2 = 0, O0 = 4, O1 = 0;
    static __device__ __forceinline__ void run(const T* q, const T* ge, const T*, T* gq, T*) { quat_to_eazyz_bwd(q, ge, gq); }
};
template <typename T> struct MatToEazyzFwd {   // lie_tools.py:178-180, fused (the quaternion never leaves registers)
    static constexpr int I0 = 9, I1 = 0, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T*, const T*, T* e, T*) {
        T q[4];
        mat_to_quat_fwd(R, q);
        quat_to_eazyz_fwd(q, e);
    }
};
template <typename T> struct MatToEazyzBwd {
    static constexpr int I0 = 9, I1 = 3, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* R, const T* ge, const T*, T* gR, T*) {
        T q[4], gq[4];
        mat_to_quat_fwd(R, q);
        quat_to_eazyz_bwd(q, ge, gq);
        mat_to_quat_bwd(R, gq, gR);
    }
};
template <typename T> struct S2S1Fwd {   // lie_tools.py:67-78
    static constexpr int I0 = 3, I1 = 2, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* u, const T* cs, const T*, T* R, T*) {
        axis_angle_matrix(u, cs[1], T(1) - cs[0], R);
    }
};
template <typename T> struct S2S1Bwd {
    static constexpr int I0 = 3, I1 = 2, I2 = 9, O0 = 3, O1 = 2;
    static __device__ __forceinline__ void run(const T* u, const T* cs, const T* G, T* gu, T* gcs) {
        T gs, gw;
        axis_angle_matrix_bwd(u, cs[1], T(1) - cs[0], G, gu, &gs, &gw);
        gcs[0] = -gw;
        gcs[1] = gs;
    }
};

template <typename T> struct S2S2Fwd {   // lie_tools.py:81-89
    static constexpr int I0 = 3, I1 = 3, I2 = 0, O0 = 9, O1 = 0;
    static __device__ __forceinline__ void run(const T* v1, const T* v2, const T*, T* R, T*) { s2s2_fwd(v1, v2, R); }
};
template <typename T> struct S2S2Bwd {
    static constexpr int I0 = 3, I1 = 3, I2 = 9, O0 = 3, O1 = 3;
    static __device__ __forceinline__ void run(const T* v1, const T* v2, const T* G, T* gv1, T* gv2) { s2s2_bwd(v1, v2, G, gv1, gv2); }
};
template <typename T> struct VecToEazyzFwd {   // lie_tools.py:92-97
    static constexpr int I0 = 3, I1 = 0, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* v, const T*, const T*, T* e, T*) {
        const T pi = T(3.14159265358979323846);
        e[0] = Sc<T>::tanh(v[0]) * pi;
        e[1] = Sc<T>::tanh(v[1]) * (pi / 2) + (pi / 2);
        e[2] = Sc<T>::tanh(v[2]) * pi;
    }
};
template <typename T> struct VecToEazyzBwd {
    static constexpr int I0 = 3, I1 = 3, I2 = 0, O0 = 3, O1 = 0;
    static __device__ __forceinline__ void run(const T* v, const T* g, const T*, T* gv, T*) {
        const T pi = T(3.14159265358979323846);
        const T t0 = Sc<T>::tanh(v[0]), t1 = Sc<T>::tanh(v[1]), t2 = Sc<T>::tanh(v[2]);
        gv[0] = g[0] * pi * (T(1) - t0 * t0);
        gv[1] = g[1] * (pi / 2) * (T(1) - t1 * t1);
        gv[2] = g[2] * pi * (T(1) - t2 * t2);
    }
};

// ------------------------------------------------------------------ SO(2)-subgroup equivariance distance
// EquivarianceLoss.forward, losses/equivariance_loss.py:27-36: g = s2s1rodrigues(e_x, (cos th, sin th)) (the rotation by th about
// the x axis), enc_rot = g . R, diff = || enc_rot - R2 ||_F^2 with R the encoding of an image and R2 the encoding of the image
// rotated by th.  One thread per sample; the residual d = g R - R2 is written next to diff for the backward:
//   g_R = 2 g_diff g^T d,   g_R2 = -2 g_diff d      (th is a random draw: no gradient).
template <typename T>
__device__ __forceinline__ void x_rotation(T th, T (&g)[9]) {
    T sn, cs;
    Sc<T>::sincos(th, &sn, &cs);
    const T ex[3] = {T(1), T(0), T(0)};
    axis_angle_matrix(ex, sn, T(1) - cs, g);          // exactly the arithmetic of s2s1rodrigues on (e_x, (cos, sin))
}
template <typename T> struct EquivSqDistFwd {
    static constexpr int I0 = 1, I1 = 9, I2 = 9, O0 = 1, O1 = 9;
    static __device__ __forceinline__ void run(const T* th, const T* R, const T* R2, T* diff, T* d) {
        T g[9];
        x_rotation(th[0], g);
        T acc = T(0);
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const T e = Sc<T>::fma(g[r * 3], R[c], Sc<T>::fma(g[r * 3 + 1], R[3 + c], g[r * 3 + 2] * R[6 + c])) - R2[r * 3 + c];
                d[r * 3 + c] = e;
                acc = Sc<T>::fma(e, e, acc);
            }
        diff[0] = acc;
    }
};
template <typename T> struct EquivSqDistBwd {
    static constexpr int I0 = 1, I1 = 9, I2 = 1, O0 = 9, O1 = 9;
    static __device__ __forceinline__ void run(const T* th, const T* d, const T* gdiff, T* gR, T* gR2) {
        T g[9];
        x_rotation(th[0], g);
        const T k = T(2) * gdiff[0];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                gR[r * 3 + c] = k * Sc<T>::fma(g[r], d[c], Sc<T>::fma(g[3 + r], d[3 + c], g[6 + r] * d[6 + c]));
                gR2[r * 3 + c] = -k * d[r * 3 + c];
            }
    }
};

// ------------------------------------------------------------------ leading-axis sum (grad reduction over n)
template <typename T>
__global__ void sum_leading_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, int64_t inner) {
    const int64_t j = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (j >= inner) return;
    T acc = 0;
    for (int64_t i = 0; i < n; ++i) acc += in[i * inner + j];
    out[j] = acc;
}

// ------------------------------------------------------------------ log-sum-exp over the leading axis
// utils.logsumexp (utils.py:4-26) as VAE.log_likelihood uses it (vae.py:164-171: importance weights over the n samples
// of a datapoint).  in (n, inner) -> out (inner).  A 256-thread CTA owns 32 consecutive columns (coalesced rows) and
// splits the n rows eight ways; every thread keeps a running (max, sum of exp) pair in ONE pass over its rows, the
// eight pairs of a column are merged through shared memory.  An all -inf column gives -inf, +inf gives +inf, NaN
// propagates -- torch semantics.
template <typename T>
__device__ __forceinline__ void lse_merge(T& m, T& s, T m2, T s2) {
    if (m2 > m) { s = s * Sc<T>::exp(m - m2) + s2; m = m2; }
    else if (m2 == m) s += s2;                                    // also covers m = m2 = +-inf without inf - inf
    else s += s2 * Sc<T>::exp(m2 - m);
}
template <typename T>
__global__ void __launch_bounds__(256)
logsumexp_leading_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, int64_t inner) {
    __shared__ T sm_m[8][33], sm_s[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t j = int64_t(blockIdx.x) * 32 + cx;
    T m = -HUGE_VAL, s = T(0);
    bool bad = false;                                             // a NaN anywhere in the column
    if (j < inner)
        for (int64_t i = ry; i < n; i += 8) {
            const T x = in[i * inner + j];
            bad |= (x != x);
            lse_merge(m, s, x, T(1));
        }
    sm_m[ry][cx] = bad ? T(NAN) : m;
    sm_s[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && j < inner) {
        T mm = sm_m[0][cx], ss = sm_s[0][cx];
        bool nan = (mm != mm);
#pragma unroll
        for (int r = 1; r < 8; ++r) {
            const T m2 = sm_m[r][cx];
            nan |= (m2 != m2);
            if (!nan) lse_merge(mm, ss, m2, sm_s[r][cx]);
        }
        out[j] = nan ? T(NAN) : ((mm == T(HUGE_VAL) || mm == -T(HUGE_VAL)) ? mm : mm + Sc<T>::log(ss));
    }
}
// g_in[i,j] = g_out[j] * exp(in[i,j] - out[j])   (softmax weights over the leading axis)
template <typename T>
__global__ void __launch_bounds__(256)
logsumexp_leading_bwd_kernel(const T* __restrict__ in, const T* __restrict__ out, const T* __restrict__ gout,
                             T* __restrict__ gin, int64_t total, int64_t inner) {
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int64_t j = idx % inner;
    const T o = out[j];
    const T w = (o == -T(HUGE_VAL)) ? T(0) : Sc<T>::exp(in[idx] - o);
    gin[idx] = gout[j] * w;
}

}  // namespace lv

// ====================================================================== C ABI
#define LV_ST(s) reinterpret_cast<cudaStream_t>(s)
#define LV_UNARY(NAME, OP)                                                                                     \
    extern "C" int lv_##NAME##_f32(const float* a, float* o, int64_t n, void* st) {                            \
        return lv::launch_rows<float, lv::OP<float>>(a, nullptr, nullptr, o, nullptr, n, LV_ST(st), #NAME);    \
    }                                                                                                          \
    extern "C" int lv_##NAME##_f64(const double* a, double* o, int64_t n, void* st) {                          \
        return lv::launch_rows<double, lv::OP<double>>(a, nullptr, nullptr, o, nullptr, n, LV_ST(st), #NAME);  \
    }
#define LV_BINARY(NAME, OP)                                                                                    \
    extern "C" int lv_##NAME##_f32(const float* a, const float* b, float* o, int64_t n, void* st) {            \
        return lv::launch_rows<float, lv::OP<float>>(a, b, nullptr, o, nullptr, n, LV_ST(st), #NAME);          \
    }                                                                                                          \
    extern "C" int lv_##NAME##_f64(const double* a, const double* b, double* o, int64_t n, void* st) {         \
        return lv::launch_rows<double, lv::OP<double>>(a, b, nullptr, o, nullptr, n, LV_ST(st), #NAME);        \
    }
#define LV_TERNARY2(NAME, OP)                                                                                  \
    extern "C" int lv_##NAME##_f32(const float* a, const float* b, const float* c, float* o, float* p,         \
                                   int64_t n, void* st) {                                                      \
        return lv::launch_rows<float, lv::OP<float>>(a, b, c, o, p, n, LV_ST(st), #NAME);                      \
    }                                                                                                          \
    extern "C" int lv_##NAME##_f64(const double* a, const double* b, const double* c, double* o, double* p,    \
                                   int64_t n, void* st) {                                                      \
        return lv::launch_rows<double, lv::OP<double>>(a, b, c, o, p, n, LV_ST(st), #NAME);                    \
    }

LV_UNARY(hat_fwd, HatFwd)
LV_UNARY(hat_bwd, HatBwd)
LV_UNARY(vee_fwd, VeeFwd)
LV_UNARY(vee_bwd, VeeBwd)
LV_UNARY(rodrigues_fwd, RodriguesFwd)
LV_BINARY(rodrigues_bwd, RodriguesBwd)
LV_UNARY(log_map_fwd, LogMapFwd)
LV_BINARY(log_map_bwd, LogMapBwd)
LV_UNARY(quat_to_mat_fwd, QuatToMatFwd)
LV_BINARY(quat_to_mat_bwd, QuatToMatBwd)
LV_UNARY(mat_to_quat_fwd, MatToQuatFwd)
LV_BINARY(mat_to_quat_bwd, MatToQuatBwd)
LV_UNARY(quat_to_eazyz_fwd, QuatToEazyzFwd)
LV_BINARY(quat_to_eazyz_bwd, QuatToEazyzBwd)
LV_UNARY(mat_to_eazyz_fwd, MatToEazyzFwd)
LV_BINARY(mat_to_eazyz_bwd, MatToEazyzBwd)
LV_BINARY(s2s1_rodrigues_fwd, S2S1Fwd)
LV_TERNARY2(s2s1_rodrigues_bwd, S2S1Bwd)
LV_BINARY(s2s2_gram_schmidt_fwd, S2S2Fwd)
LV_TERNARY2(s2s2_gram_schmidt_bwd, S2S2Bwd)
LV_UNARY(vector_to_eazyz_fwd, VecToEazyzFwd)
LV_BINARY(vector_to_eazyz_bwd, VecToEazyzBwd)
LV_TERNARY2(equivariance_sqdist_fwd, EquivSqDistFwd)
LV_TERNARY2(equivariance_sqdist_bwd, EquivSqDistBwd)

template <typename T>
static int sum_leading(const T* in, T* out, int64_t n, int64_t inner, void* st) {
    if (n < 0 || inner < 0 || (n > 0 && inner > 0 && (!in || !out))) {
        lv::set_error("sum_leading: bad arguments");
        return LV_ERR_ARG;
    }
    if (inner == 0) return LV_OK;
    const int threads = 256;
    const int64_t blocks = (inner + threads - 1) / threads;
    lv::sum_leading_kernel<T><<<unsigned(blocks), threads, 0, LV_ST(st)>>>(in, out, n, inner);
    return lv::check_launch("sum_leading");
}
template <typename T>
static int logsumexp_leading(const T* in, T* out, int64_t n, int64_t inner, void* st) {
    if (n <= 0 || inner < 0 || (inner > 0 && (!in || !out))) { lv::set_error("logsumexp_leading: bad arguments"); return LV_ERR_ARG; }
    if (inner == 0) return LV_OK;
    const int64_t blocks = (inner + 31) / 32;
    if (blocks > 0x7fffffffLL) { lv::set_error("logsumexp_leading: too many columns"); return LV_ERR_ARG; }
    lv::logsumexp_leading_kernel<T><<<unsigned(blocks), 256, 0, LV_ST(st)>>>(in, out, n, inner);
    return lv::check_launch("logsumexp_leading");
}
template <typename T>
static int logsumexp_leading_bwd(const T* in, const T* out, const T* gout, T* gin, int64_t n, int64_t inner, void* st) {
    if (n <= 0 || inner < 0 || (inner > 0 && (!in || !out || !gout || !gin))) { lv::set_error("logsumexp_leading_bwd: bad arguments"); return LV_ERR_ARG; }
    const int64_t total = n * inner;
    if (total == 0) return LV_OK;
    const int64_t blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) { lv::set_error("logsumexp_leading_bwd: too many elements"); return LV_ERR_ARG; }
    lv::logsumexp_leading_bwd_kernel<T><<<unsigned(blocks), 256, 0, LV_ST(st)>>>(in, out, gout, gin, total, inner);
    return lv::check_launch("logsumexp_leading_bwd");
}
extern "C" int lv_logsumexp_leading_fwd_f32(const float* in, float* out, int64_t n, int64_t inner, void* st) { return logsumexp_leading<float>(in, out, n, inner, st); }
extern "C" int lv_logsumexp_leading_fwd_f64(const double* in, double* out, int64_t n, int64_t inner, void* st) { return logsumexp_leading<double>(in, out, n, inner, st); }
extern "C" int lv_logsumexp_leading_bwd_f32(const float* in, const float* out, const float* gout, float* gin, int64_t n, int64_t inner, void* st) {
    return logsumexp_leading_bwd<float>(in, out, gout, gin, n, inner, st);
}
extern "C" int lv_logsumexp_leading_bwd_f64(const double* in, const double* out, const double* gout, double* gin, int64_t n, int64_t inner, void* st) {
    return logsumexp_leading_bwd<double>(in, out, gout, gin, n, inner, st);
}

extern "C" int lv_sum_leading_f32(const float* in, float* out, int64_t n, int64_t inner, void* st) {
    return sum_leading<float>(in, out, n, inner, st);
}
extern "C" int lv_sum_leading_f64(const double* in, double* out, int64_t n, int64_t inner, void* st) {
    return sum_leading<double>(in, out, n, inner, st);
}
