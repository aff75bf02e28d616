// Block-diagonal Wigner-D action on a harmonic spectrum, forward and backward (sm_100a, FP32).
//
// Replaces the reference's per-degree Python loop (lie_tools.py:226-253): for every degree it
// builds three dense X(angle) matrices (lie_tools.py:195-208), multiplies
// D^l = X(a) J X(b) J X(c) with four dense batched GEMMs (lie_tools.py:221) and bmm's D^l with
// the degree-l rows of the spectrum.  Here D^l is never formed.  The chain is applied right to
// left on the vector:
//      y_l = X(a) ( J ( X(b) ( J ( X(c) s_l ) ) ) )
// X(phi) is a set of independent 2x2 rotations of the pairs (i, 2l-i) by (l-i)*phi, and J_l is
// ~25% dense with compile-time coefficients (wigner_gen.cuh, one immediate-operand FMA per
// non-zero).  One thread owns one (sample, channel) column; the degree-l vector (<= 17 floats)
// lives in registers.  cos/sin(m*angle), m = 1..8, are computed once per sample by three threads
// (sincosf + angle-addition recurrence) and shared through smem.  The CTA's (S, M, C) output tile
// is contiguous in global memory: it is assembled in shared memory and moved with 128-bit
// coalesced accesses.  The channel count is a template parameter for the common cases so that
// every tile / spectrum access is base + immediate (no per-access integer math); CT = 0 is the
// run-time-C fallback.
//
// Backward (hand-derived; G = d/dphi X(phi) X(phi)^-1 is the pair generator
// (G w)_i = (l-i) w_{2l-i}):
//      h4 = X(a)^T g,  h3 = J h4,  h2 = X(b)^T h3,  h1 = J h2,  g_s = X(c)^T h1
//      g_a = <h4, G w4>,  g_b = <h2, G w2>,  g_c = <g_s, G s>,   w2 = J X(c) s, w4 = J X(b) w2
// so only w2 and w4 are recomputed from the spectrum; nothing is saved by the forward.
// For a shared spectrum (ActionNet.item_rep, decoders.py:53) the per-sample g_s are summed over
// the batch deterministically: rows of the tile -> per-CTA accumulator in smem (persistent CTAs)
// -> one partial per CTA -> a second tiny kernel.  No atomics anywhere.
//
// transpose=True (lie_tools.py:249-250): D^T = X(-c) J X(-b) J X(-a), i.e. the same kernels on
// the angles (-c, -b, -a), with the angle gradients mapped back.
#include "common.cuh"
#include "wigner_gen.cuh"

namespace lv {

constexpr int WG_LMAX = wg::kGenLmax;
constexpr int WG_TRIG_STRIDE = 52;   // 3 angles x 8 x (cos,sin) = 48 floats, padded: 16B-aligned rows, 4 samples on distinct bank quads
constexpr int WG_MAX_THREADS = 256;

using wg::jmul;

// pair rotations: x <- X(phi) x          (lie_tools.py:195-208: X[i,i]=cos((l-i)phi), X[i,2l-i]=sin((l-i)phi))
// cs[m-1] = (cos m phi, sin m phi); two frequencies per 128-bit LDS.  TRANSPOSED: X(phi)^T = X(-phi).
template <int L, bool TRANSPOSED>
__device__ __forceinline__ void xrot(float (&x)[2 * L + 1], const float2* __restrict__ cs) {
    const float4* cs4 = reinterpret_cast<const float4*>(cs);
#pragma unroll
    for (int p = 0; p < (L + 1) / 2; ++p) {
        float4 t = cs4[p];
        if (TRANSPOSED) { t.y = -t.y; t.w = -t.w; }
        {
            const int m = 2 * p + 1;
            const float a = x[L - m], b = x[L + m];
            x[L - m] = fmaf(t.x, a, t.y * b);
            x[L + m] = fmaf(t.x, b, -(t.y * a));
        }
        if (2 * p + 2 <= L) {
            const int m = 2 * p + 2;
            const float a = x[L - m], b = x[L + m];
            x[L - m] = fmaf(t.z, a, t.w * b);
            x[L + m] = fmaf(t.z, b, -(t.w * a));
        }
    }
}
// <h, G w> = sum_m m (h[l-m] w[l+m] - h[l+m] w[l-m])
template <int L>
__device__ __forceinline__ float gdot(const float (&h)[2 * L + 1], const float (&w)[2 * L + 1]) {
    float acc = 0.f;
#pragma unroll
    for (int m = 1; m <= L; ++m) acc = fmaf(float(m), fmaf(h[L - m], w[L + m], -(h[L + m] * w[L - m])), acc);
    return acc;
}

// cos/sin(m * angle), m = 1..WG_LMAX, for the CTA's samples.  trig[s][a][m-1] = (cos, sin);
// a = 0,1,2 are the *effective* first/second/third angles (transpose: (-c,-b,-a)).
__device__ __forceinline__ void stage_trig(float* __restrict__ s_trig, const float* __restrict__ angles, int64_t n0,
                                           int rows, int transpose) {
    for (int j = threadIdx.x; j < rows * 3; j += blockDim.x) {
        const int s = j / 3, a = j - 3 * s;
        const float phi = transpose ? -__ldg(angles + (n0 + s) * 3 + (2 - a)) : __ldg(angles + n0 * 3 + j);
        float s1, c1;
        sincosf(phi, &s1, &c1);
        float2* dst = reinterpret_cast<float2*>(s_trig + s * WG_TRIG_STRIDE + a * (2 * WG_LMAX));
        float cm = c1, sm = s1;
#pragma unroll
        for (int m = 1; m <= WG_LMAX; ++m) {
            dst[m - 1] = make_float2(cm, sm);
            const float cn = fmaf(cm, c1, -(sm * s1));
            sm = fmaf(sm, c1, cm * s1);
            cm = cn;
        }
    }
}

template <int L, bool GLOBAL_SRC>
__device__ __forceinline__ void load_col(float (&x)[2 * L + 1], const float* __restrict__ src, int stride) {
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = GLOBAL_SRC ? __ldg(src + i * stride) : src[i * stride];
}

template <int L, bool GLOBAL_SRC>
__device__ __forceinline__ void degree_fwd(const float* src, float* dst, int C, const float2* __restrict__ tg) {
    float x[2 * L + 1], y[2 * L + 1];
    load_col<L, GLOBAL_SRC>(x, src, C);
    xrot<L, false>(x, tg + 2 * WG_LMAX);
    jmul<L>(x, y);
    xrot<L, false>(y, tg + WG_LMAX);
    jmul<L>(y, x);
    xrot<L, false>(x, tg);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) dst[i * C] = x[i];
}

// g (smem tile column): upstream gradient in, spectrum gradient out (in place).
template <int L, bool GLOBAL_SRC>
__device__ __forceinline__ void degree_bwd(const float* src, float* g, int C, const float2* __restrict__ tg,
                                           float& ga, float& gb, float& gc) {
    float x[2 * L + 1], y[2 * L + 1], w2[2 * L + 1];
    load_col<L, GLOBAL_SRC>(x, src, C);
    xrot<L, false>(x, tg + 2 * WG_LMAX);
    jmul<L>(x, w2);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = w2[i];
    xrot<L, false>(y, tg + WG_LMAX);
    jmul<L>(y, x);                            // x = w4
    load_col<L, false>(y, g, C);              // y = g
    xrot<L, true>(y, tg);                     // h4
    ga += gdot<L>(y, x);
    jmul<L>(y, x);                            // x = h3
    xrot<L, true>(x, tg + WG_LMAX);           // h2
    gb += gdot<L>(x, w2);
    jmul<L>(x, y);                            // y = h1
    xrot<L, true>(y, tg + 2 * WG_LMAX);       // g_s
    load_col<L, GLOBAL_SRC>(x, src, C);
    gc += gdot<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) g[i * C] = y[i];
}

#define WG_SWITCH(l, CALL)                      \
    switch (l) {                                \
        case 0: { constexpr int L = 0; CALL; } break; \
        case 1: { constexpr int L = 1; CALL; } break; \
        case 2: { constexpr int L = 2; CALL; } break; \
        case 3: { constexpr int L = 3; CALL; } break; \
        case 4: { constexpr int L = 4; CALL; } break; \
        case 5: { constexpr int L = 5; CALL; } break; \
        case 6: { constexpr int L = 6; CALL; } break; \
        case 7: { constexpr int L = 7; CALL; } break; \
        default: { constexpr int L = 8; CALL; } break; \
    }

__host__ __device__ inline int align4i(int x) { return (x + 3) & ~3; }

// ------------------------------------------------------------------ forward
// SHARED: spectrum is (M,C), the same for every sample (stride-0 expand in the reference).
// CT: compile-time channel count (0 = use the run-time argument).
template <bool SHARED, int CT>
__global__ void __launch_bounds__(WG_MAX_THREADS)
wigner_fwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, float* __restrict__ out,
                  int64_t N, int lmin, int lmax, int Crt, int S, int transpose) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT > 0 ? CT : Crt;
    const int M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    const int MC = M * C;
    float* tile = smem;
    float* s_trig = smem + align4i(S * MC);
    const int64_t n0 = int64_t(blockIdx.x) * S;
    const int rows = int(min(int64_t(S), N - n0));
    stage_trig(s_trig, angles, n0, rows, transpose);
    if (!SHARED) tile_g2s(tile, spectrum + n0 * MC, rows * MC);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    if (s < rows) {
        const float2* tg = reinterpret_cast<const float2*>(s_trig + s * WG_TRIG_STRIDE);
        float* trow = tile + s * MC + c;
        // shared spectrum: 3 KB read by every thread of every CTA -> stays L1-resident (the kernel streams
        // nothing else through L1: outputs leave through smem), so it is read in place with LDG.
        const float* srow = SHARED ? spectrum + c : trow;
        int off = 0;
        for (int l = lmin; l <= lmax; ++l) {
            WG_SWITCH(l, (degree_fwd<L, SHARED>(srow + off, trow + off, C, tg)));
            off += (2 * l + 1) * C;
        }
    }
    __syncthreads();
    tile_s2g(out + n0 * MC, tile, rows * MC);
}

// ------------------------------------------------------------------ backward
// Persistent CTAs over sample tiles.  workspace (SHARED only): [gridDim.x][MC] partial sums.
template <bool SHARED, int CT>
__global__ void __launch_bounds__(WG_MAX_THREADS)
wigner_bwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
                  float* __restrict__ gangles, float* __restrict__ gspectrum, float* __restrict__ partial,
                  int64_t N, int lmin, int lmax, int Crt, int S, int transpose, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT > 0 ? CT : Crt;
    const int M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    const int MC = M * C;
    float* tile = smem;
    float* s_trig = tile + align4i(S * MC);
    float* s_gp = s_trig + S * WG_TRIG_STRIDE;          // [S*C][3] per-thread angle-gradient parts
    float* s_acc = s_gp + align4i(S * C * 3);           // [MC] (SHARED) batch-sum of the spectrum gradient
    float* s_item = s_acc + align4i(MC);                // [MC] (SHARED) the spectrum itself
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    if (SHARED)
        for (int o = t; o < MC; o += blockDim.x) { s_acc[o] = 0.f; s_item[o] = __ldg(spectrum + o); }
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
        const int64_t n0 = tile_idx * S;
        const int rows = int(min(int64_t(S), N - n0));
        stage_trig(s_trig, angles, n0, rows, transpose);
        tile_g2s(tile, gout + n0 * MC, rows * MC);
        tile_async_wait();
        __syncthreads();
        if (s < rows) {
            const float2* tg = reinterpret_cast<const float2*>(s_trig + s * WG_TRIG_STRIDE);
            float* trow = tile + s * MC + c;
            const float* srow = SHARED ? s_item + c : spectrum + (n0 + s) * MC + c;
            float ga = 0.f, gb = 0.f, gc = 0.f;
            int off = 0;
            for (int l = lmin; l <= lmax; ++l) {
                WG_SWITCH(l, (degree_bwd<L, !SHARED>(srow + off, trow + off, C, tg, ga, gb, gc)));
                off += (2 * l + 1) * C;
            }
            // effective angles (a',b',c') = transpose ? (-c,-b,-a) : (a,b,c)
            s_gp[t * 3 + 0] = transpose ? -gc : ga;
            s_gp[t * 3 + 1] = transpose ? -gb : gb;
            s_gp[t * 3 + 2] = transpose ? -ga : gc;
        }
        __syncthreads();
        if (SHARED) {
            for (int o = t; o < MC; o += blockDim.x) {
                float a0 = 0.f, a1 = 0.f;
                int r = 0;
                for (; r + 1 < rows; r += 2) { a0 += tile[r * MC + o]; a1 += tile[(r + 1) * MC + o]; }
                if (r < rows) a0 += tile[r * MC + o];
                s_acc[o] += a0 + a1;
            }
        } else {
            tile_s2g(gspectrum + n0 * MC, tile, rows * MC);
        }
        for (int j = t; j < rows * 3; j += blockDim.x) {
            const int ss = j / 3, a = j - 3 * ss;
            float acc = 0.f;
            for (int cc = 0; cc < C; ++cc) acc += s_gp[(ss * C + cc) * 3 + a];
            gangles[n0 * 3 + j] = acc;
        }
        __syncthreads();
    }
    if (SHARED)
        for (int o = t; o < MC; o += blockDim.x) partial[int64_t(blockIdx.x) * MC + o] = s_acc[o];
}

// partial [nblk][MC] -> out[MC]; block = (32, 8): 32 consecutive columns, 8-way split over rows.
__global__ void __launch_bounds__(256)
wigner_reduce_partials(const float* __restrict__ partial, float* __restrict__ out, int nblk, int MC) {
    __shared__ float red[8][33];
    const int o = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (o < MC)
        for (int b = threadIdx.y; b < nblk; b += 8) acc += partial[int64_t(b) * MC + o];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && o < MC) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) a += red[y][threadIdx.x];
        out[o] = a;
    }
}

// ------------------------------------------------------------------ host-side launch geometry
struct DevInfo { int sms; int smem_optin; bool ok; };
static DevInfo dev_info() {
    static DevInfo cache[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return {0, 0, false};
    if (!cache[dev].ok) {
        int sms = 0, smem = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return {0, 0, false};
        if (cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return {0, 0, false};
        cache[dev] = {sms, smem, true};
    }
    return cache[dev];
}

struct WgGeom { int M, MC, S, threads; size_t smem_fwd, smem_bwd; int64_t ntiles; int grid_bwd; };

// S = samples per CTA: about 192 threads, tile <= ~52 KB so that four CTAs share an SM.
static int wigner_geometry(const char* name, int64_t N, int lmin, int lmax, int C, bool shared, WgGeom& g) {
    if (N < 0 || C <= 0 || lmin < 0 || lmax < lmin) { set_error("%s: bad sizes (N=%lld, C=%d, degrees %d..%d)", name, (long long)N, C, lmin, lmax); return LV_ERR_ARG; }
    if (lmax > WG_LMAX) { set_error("%s: degree %d > %d is not supported by the unrolled kernels", name, lmax, WG_LMAX); return LV_ERR_UNSUPPORTED; }
    if (C > WG_MAX_THREADS) { set_error("%s: more than %d channels unsupported", name, WG_MAX_THREADS); return LV_ERR_UNSUPPORTED; }
    DevInfo di = dev_info();
    if (!di.ok) { set_error("%s: cannot query the CUDA device", name); return int(cudaErrorInvalidDevice); }
    g.M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    g.MC = g.M * C;
    int S = 192 / C;
    if (S < 1) S = 1;
    const int by_smem = (52 * 1024) / (g.MC * 4);
    if (S > by_smem) S = by_smem;
    if (S < 1) S = 1;
    if (S > 1 && (S & 1)) S -= 1;     // even S keeps tile starts 16B-aligned when MC is even
    g.S = S;
    g.threads = ((S * C + 31) / 32) * 32;
    g.smem_fwd = size_t(align4i(S * g.MC) + S * WG_TRIG_STRIDE) * 4;
    g.smem_bwd = size_t(align4i(S * g.MC) + S * WG_TRIG_STRIDE + align4i(S * C * 3) + (shared ? 2 * align4i(g.MC) : 0)) * 4;
    if (g.smem_bwd > size_t(di.smem_optin)) { set_error("%s: spectrum row of %d floats does not fit shared memory", name, g.MC); return LV_ERR_UNSUPPORTED; }
    g.ntiles = (N + S - 1) / S;
    if (g.ntiles > 0x7fffffffLL) { set_error("%s: too many samples", name); return LV_ERR_ARG; }
    const int64_t cap = int64_t(di.sms) * 4;
    g.grid_bwd = int(g.ntiles < cap ? g.ntiles : cap);
    if (g.grid_bwd < 1) g.grid_bwd = 1;
    return LV_OK;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return LV_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return int(e); }
    return LV_OK;
}

template <bool SHARED, int CT>
static int launch_fwd(const WgGeom& g, const float* angles, const float* spectrum, float* out, int64_t N, int lmin,
                      int lmax, int C, int transpose, cudaStream_t st) {
    int rc = opt_in_smem(wigner_fwd_kernel<SHARED, CT>, g.smem_fwd);
    if (rc) return rc;
    wigner_fwd_kernel<SHARED, CT><<<unsigned(g.ntiles), g.threads, g.smem_fwd, st>>>(angles, spectrum, out, N, lmin, lmax, C, g.S, transpose);
    return check_launch("wigner_apply_fwd");
}

template <bool SHARED, int CT>
static int launch_bwd(const WgGeom& g, const float* angles, const float* spectrum, const float* gout, float* gangles,
                      float* gspectrum, float* workspace, int64_t N, int lmin, int lmax, int C, int transpose,
                      cudaStream_t st) {
    int rc = opt_in_smem(wigner_bwd_kernel<SHARED, CT>, g.smem_bwd);
    if (rc) return rc;
    wigner_bwd_kernel<SHARED, CT><<<g.grid_bwd, g.threads, g.smem_bwd, st>>>(angles, spectrum, gout, gangles, SHARED ? nullptr : gspectrum,
                                                                             SHARED ? workspace : nullptr, N, lmin, lmax, C, g.S, transpose, g.ntiles);
    return check_launch("wigner_apply_bwd");
}

}  // namespace lv

// channel counts with a compile-time specialisation: 10 = ActionNet default (decoders.py:11, main.py:168)
#define WG_DISPATCH_C(C, SHARED, FN, ...) ((C) == 10 ? FN<SHARED, 10>(__VA_ARGS__) : FN<SHARED, 0>(__VA_ARGS__))

// ====================================================================== C ABI
extern "C" int64_t lv_wigner_bwd_workspace_floats(int64_t N, int lmin, int lmax, int C) {
    lv::WgGeom g;
    if (lv::wigner_geometry("wigner_bwd_workspace", N, lmin, lmax, C, true, g) != LV_OK) return -1;
    return int64_t(g.grid_bwd) * g.MC;
}

extern "C" int lv_wigner_apply_fwd_f32(const float* angles, const float* spectrum, float* out, int64_t N, int lmin,
                                       int lmax, int C, int shared_spectrum, int transpose, void* stream) {
    lv::WgGeom g;
    int rc = lv::wigner_geometry("wigner_apply_fwd", N, lmin, lmax, C, shared_spectrum != 0, g);
    if (rc) return rc;
    if (N == 0) return LV_OK;
    if (!angles || !spectrum || !out) { lv::set_error("wigner_apply_fwd: null pointer"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (shared_spectrum) return WG_DISPATCH_C(C, true, lv::launch_fwd, g, angles, spectrum, out, N, lmin, lmax, C, transpose, st);
    return WG_DISPATCH_C(C, false, lv::launch_fwd, g, angles, spectrum, out, N, lmin, lmax, C, transpose, st);
}

extern "C" int lv_wigner_apply_bwd_f32(const float* angles, const float* spectrum, const float* gout, float* gangles,
                                       float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int lmin,
                                       int lmax, int C, int shared_spectrum, int transpose, void* stream) {
    lv::WgGeom g;
    int rc = lv::wigner_geometry("wigner_apply_bwd", N, lmin, lmax, C, shared_spectrum != 0, g);
    if (rc) return rc;
    if (!gspectrum) { lv::set_error("wigner_apply_bwd: null pointer"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (N == 0) {
        if (shared_spectrum) {
            cudaError_t e = cudaMemsetAsync(gspectrum, 0, size_t(g.MC) * 4, st);
            if (e != cudaSuccess) { lv::set_error("wigner_apply_bwd: memset: %s", cudaGetErrorString(e)); return int(e); }
        }
        return LV_OK;
    }
    if (!angles || !spectrum || !gout || !gangles) { lv::set_error("wigner_apply_bwd: null pointer"); return LV_ERR_ARG; }
    if (shared_spectrum) {
        const int64_t need = int64_t(g.grid_bwd) * g.MC;
        if (!workspace || workspace_floats < need) { lv::set_error("wigner_apply_bwd: workspace of %lld floats required", (long long)need); return LV_ERR_ARG; }
        rc = WG_DISPATCH_C(C, true, lv::launch_bwd, g, angles, spectrum, gout, gangles, gspectrum, workspace, N, lmin, lmax, C, transpose, st);
        if (rc) return rc;
        lv::wigner_reduce_partials<<<(g.MC + 31) / 32, dim3(32, 8), 0, st>>>(workspace, gspectrum, g.grid_bwd, g.MC);
        return lv::check_launch("wigner_reduce_partials");
    }
    return WG_DISPATCH_C(C, false, lv::launch_bwd, g, angles, spectrum, gout, gangles, gspectrum, workspace, N, lmin, lmax, C, transpose, st);
}
