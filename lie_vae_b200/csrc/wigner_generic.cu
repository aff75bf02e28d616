// Block-diagonal Wigner-D action for ANY degree range and FP32 / FP64 (sm_100a): the general path behind
// block_wigner_matrix_multiply / wigner_d_matrix / ActionNet (lie_tools.py:195-253, decoders.py:47-56) for what the
// unrolled, packed kernels of wigner.cu do not cover: degrees above 8 and float64 tensors.
//
// Same algorithm as wigner.cu -- D^l is never formed, the chain  X(a) J X(b) J X(c)  is applied right to left to the
// degree-l vector of one (sample, channel) column, the backward recomputes w2 = J X(c) s and w4 = J X(b) w2 -- but with
// run-time loops: J_l comes from a caller-owned dense table in global memory (all threads read the same element: one
// broadcast transaction; structural zeros, 75 % of J, are skipped warp-uniformly), the three work vectors of a thread
// live in shared memory ([vector][element][thread]: conflict-free), cos/sin(m phi) follow from the angle-addition
// recurrence as m runs, so nothing scales with the degree except the loops.
// Outputs of the backward are per column: g_spectrum (N,M,C) and the angle-gradient parts (N,C,3); the host wrapper
// sums over channels / over the batch for a shared spectrum (deterministic, no atomics).
#include "common.cuh"

namespace lv {

constexpr int WGEN_LMAX = 32;

__host__ __device__ inline int64_t j_offset(int l) { return int64_t(l) * (2 * l - 1) * (2 * l + 1) / 3; }   // sum_{k<l} (2k+1)^2

template <typename T> struct SVec {         // element i of this thread's vector
    T* p; int stride;
    __device__ __forceinline__ T& operator[](int i) const { return p[i * stride]; }
};

// x <- X(phi) x  (TR: X(phi)^T = X(-phi));  (c1, s1) = (cos phi, sin phi)
template <typename T, bool TR>
__device__ __forceinline__ void gen_xrot(const SVec<T>& x, int l, T c1, T s1) {
    if (TR) s1 = -s1;
    T cm = c1, sm = s1;
    for (int m = 1; m <= l; ++m) {
        const T a = x[l - m], b = x[l + m];
        x[l - m] = Sc<T>::fma(cm, a, sm * b);
        x[l + m] = Sc<T>::fma(cm, b, -(sm * a));
        const T cn = Sc<T>::fma(cm, c1, -(sm * s1));
        sm = Sc<T>::fma(sm, c1, cm * s1);
        cm = cn;
    }
}
// y = J_l x
template <typename T>
__device__ __forceinline__ void gen_jmul(const T* __restrict__ J, const SVec<T>& x, const SVec<T>& y, int d) {
    for (int i = 0; i < d; ++i) {
        T acc = T(0);
        const T* row = J + i * d;
        for (int j = 0; j < d; ++j) {
            const T k = __ldg(row + j);
            if (k != T(0)) acc = Sc<T>::fma(k, x[j], acc);        // warp-uniform branch
        }
        y[i] = acc;
    }
}
// <h, G w> = sum_m m (h[l-m] w[l+m] - h[l+m] w[l-m])
template <typename T>
__device__ __forceinline__ T gen_gdot(const SVec<T>& h, const SVec<T>& w, int l) {
    T acc = T(0);
    for (int m = 1; m <= l; ++m) acc = Sc<T>::fma(T(m), Sc<T>::fma(h[l - m], w[l + m], -(h[l + m] * w[l - m])), acc);
    return acc;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(128)
wigner_generic_kernel(const T* __restrict__ angles, const T* __restrict__ spectrum, const T* __restrict__ Jtab,
                      const T* __restrict__ gout, T* __restrict__ out, T* __restrict__ gang_parts, T* __restrict__ gspec,
                      int64_t N, int lmin, int lmax, int C, int shared, int transpose) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int dmax = 2 * lmax + 1, bd = blockDim.x, tid = threadIdx.x;
    const int64_t col = int64_t(blockIdx.x) * bd + tid;
    if (col >= N * C) return;
    const int64_t n = col / C;
    const int c = int(col - n * C);
    const int M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    const SVec<T> x{smem + tid, bd}, y{smem + dmax * bd + tid, bd}, w2{smem + 2 * dmax * bd + tid, bd};
    // effective angles (a,b,c); transpose: D^T = X(-c) J X(-b) J X(-a)
    T ang[3], cs[3], sn[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) ang[a] = transpose ? -angles[n * 3 + (2 - a)] : angles[n * 3 + a];
#pragma unroll
    for (int a = 0; a < 3; ++a) Sc<T>::sincos(ang[a], &sn[a], &cs[a]);
    const T* srow = spectrum + (shared ? 0 : n * int64_t(M) * C) + c;
    T ga = T(0), gb = T(0), gc = T(0);
    for (int l = lmin; l <= lmax; ++l) {
        const int d = 2 * l + 1;
        const int64_t r0 = int64_t(l * l - lmin * lmin) * C;
        const T* J = Jtab + j_offset(l);
        for (int i = 0; i < d; ++i) x[i] = srow[r0 + int64_t(i) * C];
        gen_xrot<T, false>(x, l, cs[2], sn[2]);
        if (!BWD) {
            gen_jmul(J, x, y, d);
            gen_xrot<T, false>(y, l, cs[1], sn[1]);
            gen_jmul(J, y, x, d);
            gen_xrot<T, false>(x, l, cs[0], sn[0]);
            T* o = out + n * int64_t(M) * C + r0 + c;
            for (int i = 0; i < d; ++i) o[int64_t(i) * C] = x[i];
        } else {
            gen_jmul(J, x, w2, d);
            for (int i = 0; i < d; ++i) y[i] = w2[i];
            gen_xrot<T, false>(y, l, cs[1], sn[1]);
            gen_jmul(J, y, x, d);                                   // x = w4
            const T* g = gout + n * int64_t(M) * C + r0 + c;
            for (int i = 0; i < d; ++i) y[i] = g[int64_t(i) * C];
            gen_xrot<T, true>(y, l, cs[0], sn[0]);                  // h4
            ga += gen_gdot(y, x, l);
            gen_jmul(J, y, x, d);                                   // h3
            gen_xrot<T, true>(x, l, cs[1], sn[1]);                  // h2
            gb += gen_gdot(x, w2, l);
            gen_jmul(J, x, y, d);                                   // h1
            gen_xrot<T, true>(y, l, cs[2], sn[2]);                  // g_s
            for (int i = 0; i < d; ++i) x[i] = srow[r0 + int64_t(i) * C];
            gc += gen_gdot(y, x, l);
            T* o = gspec + n * int64_t(M) * C + r0 + c;
            for (int i = 0; i < d; ++i) o[int64_t(i) * C] = y[i];
        }
    }
    if (BWD) {
        T* gp = gang_parts + col * 3;
        gp[0] = transpose ? -gc : ga;
        gp[1] = transpose ? -gb : gb;
        gp[2] = transpose ? -ga : gc;
    }
}

template <typename T, bool BWD>
static int launch_generic(const char* name, const T* angles, const T* spectrum, const T* J, const T* gout, T* out,
                          T* gang_parts, T* gspec, int64_t N, int lmin, int lmax, int C, int shared, int transpose, void* stream) {
    if (N < 0 || C <= 0 || lmin < 0 || lmax < lmin) { set_error("%s: bad sizes (N=%lld, C=%d, degrees %d..%d)", name, (long long)N, C, lmin, lmax); return LV_ERR_ARG; }
    if (lmax > WGEN_LMAX) { set_error("%s: degree %d > %d is not supported", name, lmax, WGEN_LMAX); return LV_ERR_UNSUPPORTED; }
    if (N == 0) return LV_OK;
    if (!angles || !spectrum || !J || (BWD ? (!gout || !gang_parts || !gspec) : !out)) { set_error("%s: null pointer", name); return LV_ERR_ARG; }
    const int64_t cols = N * C;
    int bd = sizeof(T) == 8 ? 64 : 128;
    const size_t per_thread = size_t(3) * (2 * lmax + 1) * sizeof(T);
    while (bd > 32 && per_thread * bd > 160 * 1024) bd >>= 1;
    const size_t smem = per_thread * bd;
    const int64_t grid = (cols + bd - 1) / bd;
    if (grid > 0x7fffffffLL) { set_error("%s: too many columns", name); return LV_ERR_ARG; }
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(wigner_generic_kernel<T, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) { set_error("%s: cudaFuncSetAttribute(smem=%zu): %s", name, smem, cudaGetErrorString(e)); return int(e); }
    }
    wigner_generic_kernel<T, BWD><<<unsigned(grid), bd, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        angles, spectrum, J, gout, out, gang_parts, gspec, N, lmin, lmax, C, shared, transpose);
    return check_launch(name);
}

}  // namespace lv

// ====================================================================== C ABI
extern "C" int lv_wigner_generic_max_degree(void) { return lv::WGEN_LMAX; }

#define LV_WGEN_ENTRY(SFX, T)                                                                                                   \
    extern "C" int lv_wigner_generic_fwd_##SFX(const T* angles, const T* spectrum, const T* jtable, T* out, int64_t N, int lmin, \
                                               int lmax, int C, int shared_spectrum, int transpose, void* stream) {             \
        return lv::launch_generic<T, false>("wigner_generic_fwd", angles, spectrum, jtable, nullptr, out, nullptr, nullptr, N,   \
                                            lmin, lmax, C, shared_spectrum, transpose, stream);                                  \
    }                                                                                                                           \
    extern "C" int lv_wigner_generic_bwd_##SFX(const T* angles, const T* spectrum, const T* jtable, const T* gout,               \
                                               T* gangle_parts, T* gspectrum, int64_t N, int lmin, int lmax, int C,             \
                                               int shared_spectrum, int transpose, void* stream) {                              \
        return lv::launch_generic<T, true>("wigner_generic_bwd", angles, spectrum, jtable, gout, nullptr, gangle_parts,          \
                                           gspectrum, N, lmin, lmax, C, shared_spectrum, transpose, stream);                     \
    }
LV_WGEN_ENTRY(f32, float)
LV_WGEN_ENTRY(f64, double)
