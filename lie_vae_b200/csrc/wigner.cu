// Block-diagonal Wigner-D action on a harmonic spectrum, forward and backward (sm_100a, FP32).
//
// Replaces the reference's per-degree Python loop (lie_tools.py:226-253): for every degree it
// builds three dense X(angle) matrices (lie_tools.py:195-208), multiplies
// D^l = X(a) J X(b) J X(c) with four dense batched GEMMs (lie_tools.py:221) and bmm's D^l with
// the degree-l rows of the spectrum.  Here D^l is never formed.  The chain is applied right to
// left on the vector:
//      y_l = X(a) ( J ( X(b) ( J ( X(c) s_l ) ) ) )
// X(phi) is a set of independent 2x2 rotations of the pairs (i, 2l-i) by (l-i)*phi, and J_l is
// ~25% dense with compile-time coefficients.  One thread owns one (sample, channel) column; the
// degree-l vector (<= 17 floats) lives in registers, packed pairwise so that every operator of the
// chain is issued as two-wide FP32 instructions (fma/mul.rn.f32x2 -> SASS FFMA2/FMUL2; layout and
// code in the generated wigner_gen.cuh): the kernels are bound by FP32 instruction issue / the FMA
// pipe, not by memory.  cos/sin(m*angle), m = 1..8, are computed once per sample (sincosf +
// angle-addition recurrence) and shared through smem in the pairing order of the packed operators.
//
// Data movement (Blackwell): the CTA's (S, M, C) tile is one contiguous span of HBM.
//   forward : columns are assembled in shared memory and the whole tile leaves with ONE TMA bulk
//             store (cp.async.bulk.global.shared::cta) issued by one thread -- no per-thread
//             copy-out loop, the LSU and the issue slots stay with the math;
//   backward: (shared spectrum, the ActionNet case) one persistent CTA per SM: math warps pull
//             32-column slices of TMA-loaded tiles from a work counter, five producer warps recycle
//             the four tile buffers (column-sum batch reduction, TMA refill, trig tables) -- see
//             wigner_bwd_ws_kernel.  (per-sample spectrum, other C / degree ranges, ragged tails:
//             wigner_bwd_kernel, where the tile lands by cp.async and the spectrum gradient
//             overwrites it in place.)
// Specialisations: the channel count (10, the ActionNet default) and the degree range (0..8 and
// 0..6) are template parameters for the common cases, so the degree loop is fully unrolled and every
// tile / spectrum access is base + immediate; <CT = 0, LT = -1> is the run-time fallback for
// any C <= 256 and any 0 <= lmin <= lmax <= 8.
//
// Backward (hand-derived).  Only the backward chain runs per column:
//      h4 = X(a)^T g,  h3 = J h4,  h2 = X(b)^T h3,  h1 = J h2,  g_s = X(c)^T h1          (g_s = D^T g: the spectrum gradient)
// The angle gradients need no forward intermediates.  X(phi) = exp(phi G_z) with the pair generator (G_z w)_i = (l-i) w_{2l-i};
// G_y = J G_z J and G_x = [G_z, G_y] complete a representation of so(3), and differentiating D = exp(a G_z) exp(b G_y) exp(c G_z)
// in the body frame gives
//      D^-1 dD/dc = G_z,    D^-1 dD/db = -sin(c) G_x + cos(c) G_y,    D^-1 dD/da = sin(b) cos(c) G_x + sin(b) sin(c) G_y + cos(b) G_z
// so d<g, D s>/d(angle) is a 3x3 combination of T_k = <g_s, G_k s>, k = x, y, z: sparse bilinear forms of the column's g_s and s
// (coefficients generated as immediates, wigner_gen.cuh).  Nothing is saved by the forward and nothing of it is recomputed.
// For a shared spectrum (ActionNet.item_rep, decoders.py:53) the per-sample g_s are summed over
// the batch reproducibly: tiles are assigned to CTAs statically and every CTA adds its tiles' rows
// in a fixed order (warp-decoupled kernel: the producer warps' register accumulators; cp.async
// kernel: per-CTA smem accumulator), then one partial row per CTA -> a second tiny kernel.  No atomics.
//
// transpose=True (lie_tools.py:249-250): D^T = X(-c) J X(-b) J X(-a), i.e. the same kernels on
// the angles (-c, -b, -a), with the angle gradients mapped back.
#include <stdlib.h>
#include "common.cuh"
#include "wigner_gen.cuh"

namespace lv {

constexpr int WG_LMAX = wg2::kGenLmax;
constexpr int WG_TRIG_STRIDE = 52;   // 3 angles x 8 x (cos,sin) = 48 floats, padded: 16B-aligned rows, 4 samples on distinct bank quads
constexpr int WG_MAX_THREADS = 256;
constexpr int WG_MAX_CTAS_PER_SM = 8;   // bound used to size the backward workspace

using wg2::PDeg;
using wg2::f32x2_t;

// Per-sample trig table: 3 angles x 4 float4 slots; slot q of an angle holds (cos m1 phi, cos m2 phi, sin m1 phi, sin m2 phi)
// for the frequency pairs (1,3) (5,7) (2,4) (6,8) -- the pairing of wigner_gen.cuh, so that a pair rotation
// (lie_tools.py:195-208: X[i,i] = cos((l-i)phi), X[i,2l-i] = sin((l-i)phi)) reads its operands with one LDS.128.
constexpr int WG_TRIG_ANGLE = 16;   // floats per angle
__device__ __forceinline__ int trig_index(int m) {      // float offset of cos(m phi) inside an angle's block; sin is +2
    return (m & 1) ? ((m - 1) >> 2) * 4 + (((m - 1) >> 1) & 1) : 8 + ((m - 2) >> 2) * 4 + (((m - 2) >> 1) & 1);
}
// cos/sin(m * phi), m = 1..WG_LMAX, by sincosf + angle-addition recurrence, written in table order
__device__ __forceinline__ void trig_fill(float* __restrict__ dst, float phi) {
    float s1, c1;
    sincosf(phi, &s1, &c1);
    float cm = c1, sm = s1;
#pragma unroll
    for (int m = 1; m <= WG_LMAX; ++m) {
        dst[trig_index(m)] = cm;
        dst[trig_index(m) + 2] = sm;
        const float cn = fmaf(cm, c1, -(sm * s1));
        sm = fmaf(sm, c1, cm * s1);
        cm = cn;
    }
}

// cos/sin(m * angle) for the CTA's samples.  trig[s][a] with a = 0,1,2 the *effective* first/second/third
// angles (transpose: (-c,-b,-a)).
__device__ __forceinline__ void stage_trig(float* __restrict__ s_trig, const float* __restrict__ angles, int64_t n0,
                                           int rows, int transpose) {
    for (int j = threadIdx.x; j < rows * 3; j += blockDim.x) {
        const int s = j / 3, a = j - 3 * s;
        const float phi = transpose ? -__ldg(angles + (n0 + s) * 3 + (2 - a)) : __ldg(angles + n0 * 3 + j);
        trig_fill(s_trig + s * WG_TRIG_STRIDE + a * WG_TRIG_ANGLE, phi);
    }
}

// y_l = X(a) J X(b) J X(c) s_l on one column; tg = the sample's trig table (angle a at +0, b at +4, c at +8 float4)
template <int L, bool GLOBAL_SRC>
__device__ __forceinline__ void degree_fwd(const float* src, float* dst, int C, const float4* __restrict__ tg) {
    using D = PDeg<L>;
    typename D::Vec x, y;
    D::template load<GLOBAL_SRC>(x, src, C);
    D::template xrot<false>(x, tg + 8);
    D::jmul(x, y);
    D::template xrot<false>(y, tg + 4);
    D::jmul(y, x);
    D::template xrot<false>(x, tg);
    D::store(x, dst, C);
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

// all degrees of one column.  LT >= 0: degrees 0..LT, fully unrolled (offsets l^2*C are constants);
// LT < 0: run-time range lmin..lmax.
template <int L, int LT, bool GLOBAL_SRC>
__device__ __forceinline__ void fwd_unrolled(const float* srow, float* trow, int C, const float4* tg) {
    degree_fwd<L, GLOBAL_SRC>(srow + L * L * C, trow + L * L * C, C, tg);
    if constexpr (L < LT) fwd_unrolled<L + 1, LT, GLOBAL_SRC>(srow, trow, C, tg);
}
template <int LT, bool GLOBAL_SRC>
__device__ __forceinline__ void fwd_degrees(const float* srow, float* trow, int C, const float4* tg, int lmin, int lmax) {
    if constexpr (LT >= 0) {
        fwd_unrolled<0, LT, GLOBAL_SRC>(srow, trow, C, tg);
    } else {
        int off = 0;
        for (int l = lmin; l <= lmax; ++l) {
            WG_SWITCH(l, (degree_fwd<L, GLOBAL_SRC>(srow + off, trow + off, C, tg)));
            off += (2 * l + 1) * C;
        }
    }
}
// Backward of one degree without the forward recompute (shared or per-sample spectrum alike): only the backward chain
// g_s = D^T g runs; the angle gradients follow from the body-frame generators of the chain D = X(a) J X(b) J X(c),
//     G_z = X'(0),  G_y = J G_z J,  G_x = [G_z, G_y]          (a representation of so(3); coefficients in wigner_gen.cuh)
//     D^-1 dD/dc = G_z,   D^-1 dD/db = -sin(c) G_x + cos(c) G_y,   D^-1 dD/da = sin(b) cos(c) G_x + sin(b) sin(c) G_y + cos(b) G_z
// i.e. d<g, D s>/d(angle) is a 3x3 combination (angle_grads_from_generators) of T_k = <g_s, G_k s>, k = x, y, z, which are
// sparse bilinear forms of the column's g_s and s.  1 440 instead of 2 140 FMA-pipe cycles per column, one spectrum load
// instead of two, no w2 kept in registers.
struct GenAcc {
    f32x2_t pz = 0ull;
    float tx = 0.f, ty = 0.f, sz = 0.f;
    __device__ __forceinline__ float tz() const { return sz + (wg2::plo(pz) + wg2::phi(pz)); }
};
template <int L, bool GLOBAL_SRC>
__device__ __forceinline__ void degree_bwd_gen(const float* src, float* g, int C, const float4* __restrict__ tg, GenAcc& acc) {
    using D = PDeg<L>;
    typename D::Vec x, y;
    D::template load<false>(y, g, C);                // y = g
    D::template xrot<true>(y, tg);                   // h4 = X(a)^T g
    D::jmul(y, x);                                   // h3
    D::template xrot<true>(x, tg + 4);               // h2
    D::jmul(x, y);                                   // h1
    D::template xrot<true>(y, tg + 8);               // g_s
    D::template load<GLOBAL_SRC>(x, src, C);         // x = s
    D::gdot(y, x, acc.pz, acc.sz);                   // T_z
    D::gxy_dots(y, x, acc.tx, acc.ty);               // T_x, T_y
    D::store(y, g, C);
}
template <int L, int LT, bool GLOBAL_SRC>
__device__ __forceinline__ void bwd_gen_unrolled(const float* srow, float* trow, int C, const float4* tg, GenAcc& acc) {
    degree_bwd_gen<L, GLOBAL_SRC>(srow + L * L * C, trow + L * L * C, C, tg, acc);
    if constexpr (L < LT) bwd_gen_unrolled<L + 1, LT, GLOBAL_SRC>(srow, trow, C, tg, acc);
}
// (T_x, T_y, T_z) and cos / sin of the second and third angle -> gradient of angle `which` (0, 1, 2)
__device__ __forceinline__ float angle_grad_from_generators(int which, float tx, float ty, float tz, float cb, float sb, float cc, float sc) {
    if (which == 0) return fmaf(sb, fmaf(cc, tx, sc * ty), cb * tz);
    if (which == 1) return fmaf(cc, ty, -(sc * tx));
    return tz;
}

template <int LT, bool GLOBAL_SRC>
__device__ __forceinline__ void bwd_degrees(const float* srow, float* trow, int C, const float4* tg, int lmin, int lmax,
                                            GenAcc& acc) {
    if constexpr (LT >= 0) {
        bwd_gen_unrolled<0, LT, GLOBAL_SRC>(srow, trow, C, tg, acc);
    } else {
        int off = 0;
        for (int l = lmin; l <= lmax; ++l) {
            WG_SWITCH(l, (degree_bwd_gen<L, GLOBAL_SRC>(srow + off, trow + off, C, tg, acc)));
            off += (2 * l + 1) * C;
        }
    }
}

__host__ __device__ inline int align4i(int x) { return (x + 3) & ~3; }

// launch bounds of the specialised kernels: 16 samples x 10 channels = 160 threads per CTA
__host__ __device__ constexpr int wg_threads(int CT) { return CT == 10 ? 160 : WG_MAX_THREADS; }

// ------------------------------------------------------------------ forward
// SHARED: spectrum is (M,C), the same for every sample (stride-0 expand in the reference).
template <bool SHARED, int CT, int LT>
__global__ void __launch_bounds__(wg_threads(CT), CT == 10 ? 4 : 1)
wigner_fwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, float* __restrict__ out,
                  int64_t N, int lmin_rt, int lmax_rt, int Crt, int S, int transpose) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT > 0 ? CT : Crt;
    const int lmin = LT >= 0 ? 0 : lmin_rt, lmax = LT >= 0 ? LT : lmax_rt;
    const int M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    const int MC = M * C;
    float* tile = smem;
    float* s_trig = smem + align4i(S * MC);
    const int64_t n0 = int64_t(blockIdx.x) * S;
    const int rows = int(min(int64_t(S), N - n0));
    if (!SHARED) tile_g2s(tile, spectrum + n0 * MC, rows * MC);
    stage_trig(s_trig, angles, n0, rows, transpose);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    if (s < rows) {
        const float4* tg = reinterpret_cast<const float4*>(s_trig + s * WG_TRIG_STRIDE);
        float* trow = tile + s * MC + c;
        // shared spectrum: 3 KB read by every thread of every CTA -> stays L1-resident (the kernel streams
        // nothing else through L1: outputs leave through smem), so it is read in place with LDG.
        const float* srow = SHARED ? spectrum + c : trow;
        fwd_degrees<LT, SHARED>(srow, trow, C, tg, lmin, lmax);
    }
    float* gdst = out + n0 * MC;
    const uint32_t bytes = uint32_t(rows) * uint32_t(MC) * 4u;
    if ((bytes & 15u) == 0 && (reinterpret_cast<uintptr_t>(gdst) & 15u) == 0) {
        // TMA bulk store: make the generic-proxy smem writes visible to the async proxy, then one thread
        // hands the whole tile to the copy engine and waits only until the engine has read it.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(gdst), "r"(uint32_t(__cvta_generic_to_shared(tile))), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        __syncthreads();
        tile_s2g(gdst, tile, rows * MC);
    }
}

// ------------------------------------------------------------------ forward fused with the reconstruction term
// VAE.log_likelihood (experiments/vae.py:164-171; n = 500 importance samples of one datapoint) needs only
// log p(x|z) = -sum_{m,c} (y - x)^2 of the decoded harmonics (VAE.recon_loss, vae.py:199-204, with the toy deconv): the
// action output y (n*B, M, C) is reduced against x (B, M, C) inside the kernel and never written.  Sample i belongs to
// datapoint i % B (the (n, B) layout of the reference).  Run-time C and degree range; forward only (evaluation path).
__global__ void __launch_bounds__(WG_MAX_THREADS, 1)
wigner_sse_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ x, float* __restrict__ out,
                  int64_t N, int64_t B, int lmin, int lmax, int C, int S, int transpose) {
    extern __shared__ __align__(16) float smem[];
    const int M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    const int MC = M * C;
    float* tile = smem;
    float* s_trig = smem + align4i(S * MC);
    float* s_part = s_trig + S * WG_TRIG_STRIDE;        // [S * C] per-column sums of squares
    const int64_t n0 = int64_t(blockIdx.x) * S;
    const int rows = int(min(int64_t(S), N - n0));
    stage_trig(s_trig, angles, n0, rows, transpose);
    __syncthreads();
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    if (s < rows) {
        const float4* tg = reinterpret_cast<const float4*>(s_trig + s * WG_TRIG_STRIDE);
        float* trow = tile + s * MC + c;
        fwd_degrees<-1, true>(spectrum + c, trow, C, tg, lmin, lmax);
        const float* xrow = x + ((n0 + s) % B) * MC + c;
        float acc = 0.f;
        for (int m = 0; m < M; ++m) {
            const float d = trow[m * C] - __ldg(xrow + m * C);
            acc = fmaf(d, d, acc);
        }
        s_part[t] = acc;
    }
    __syncthreads();
    if (t < rows) {
        float a = 0.f;
        for (int cc = 0; cc < C; ++cc) a += s_part[t * C + cc];
        out[n0 + t] = a;
    }
}

// ------------------------------------------------------------------ backward
// Persistent CTAs over sample tiles.  workspace (SHARED only): [gridDim.x][MC] partial sums.
template <bool SHARED, int CT, int LT>
__global__ void __launch_bounds__(wg_threads(CT), CT == 10 ? 3 : 1)
wigner_bwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
                  float* __restrict__ gangles, float* __restrict__ gspectrum, float* __restrict__ partial,
                  int64_t N, int lmin_rt, int lmax_rt, int Crt, int S, int transpose, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    const int C = CT > 0 ? CT : Crt;
    const int lmin = LT >= 0 ? 0 : lmin_rt, lmax = LT >= 0 ? LT : lmax_rt;
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
        tile_g2s(tile, gout + n0 * MC, rows * MC);       // asynchronous: in flight while the trig table is built
        stage_trig(s_trig, angles, n0, rows, transpose);
        tile_async_wait();
        __syncthreads();
        if (s < rows) {
            const float4* tg = reinterpret_cast<const float4*>(s_trig + s * WG_TRIG_STRIDE);
            float* trow = tile + s * MC + c;
            const float* srow = SHARED ? s_item + c : spectrum + (n0 + s) * MC + c;
            GenAcc acc;
            bwd_degrees<LT, !SHARED>(srow, trow, C, tg, lmin, lmax, acc);
            s_gp[t * 3 + 0] = acc.tx;            // (T_x, T_y, T_z) of this column
            s_gp[t * 3 + 1] = acc.ty;
            s_gp[t * 3 + 2] = acc.tz();
        }
        __syncthreads();
        if (SHARED) {
            if ((MC & 1) == 0) {
                // column sums over the tile rows, four columns per thread with 64-bit LDS (rows are 8B-aligned)
                const int nq = (MC + 3) >> 2;
                for (int q = t; q < nq; q += blockDim.x) {
                    const int o = 4 * q;
                    const bool full = o + 3 < MC;          // MC even: the tail quad holds exactly two columns
                    float2 a = make_float2(0.f, 0.f), b = make_float2(0.f, 0.f);
                    for (int r = 0; r < rows; ++r) {
                        const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + o);
                        a.x += u.x; a.y += u.y;
                        if (full) {
                            const float2 v = *reinterpret_cast<const float2*>(tile + r * MC + o + 2);
                            b.x += v.x; b.y += v.y;
                        }
                    }
                    s_acc[o] += a.x; s_acc[o + 1] += a.y;
                    if (full) { s_acc[o + 2] += b.x; s_acc[o + 3] += b.y; }
                }
            } else {
                for (int o = t; o < MC; o += blockDim.x) {
                    float a0 = 0.f;
                    for (int r = 0; r < rows; ++r) a0 += tile[r * MC + o];
                    s_acc[o] += a0;
                }
            }
        } else {
            tile_s2g(gspectrum + n0 * MC, tile, rows * MC);
        }
        for (int j = t; j < rows * 3; j += blockDim.x) {
            // T_k summed over the channels, then the body-frame relation for the effective angles (a',b',c') = transpose ?
            // (-c,-b,-a) : (a,b,c), whose cos / sin are in the trig table
            const int ss = j / 3, a = j - 3 * ss;
            float tx = 0.f, ty = 0.f, tz = 0.f;
            for (int cc = 0; cc < C; ++cc) {
                const float* gp = s_gp + (ss * C + cc) * 3;
                tx += gp[0]; ty += gp[1]; tz += gp[2];
            }
            const float* tr_s = s_trig + ss * WG_TRIG_STRIDE;
            const float gval = angle_grad_from_generators(transpose ? 2 - a : a, tx, ty, tz, tr_s[WG_TRIG_ANGLE], tr_s[WG_TRIG_ANGLE + 2],
                                                          tr_s[2 * WG_TRIG_ANGLE], tr_s[2 * WG_TRIG_ANGLE + 2]);
            gangles[n0 * 3 + j] = transpose ? -gval : gval;
        }
        __syncthreads();
    }
    if (SHARED)
        for (int o = t; o < MC; o += blockDim.x) partial[int64_t(blockIdx.x) * MC + o] = s_acc[o];
}

// mbarrier / TMA helpers: common.cuh
__device__ __forceinline__ void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ------------------------------------------------------------------ backward, shared spectrum, warp-decoupled (sm_100a)
// Same math as above, no group barriers.  One persistent 512-thread CTA per SM: 11 math warps and five producer warps,
// four (degrees 0..8) or six (0..6) 16-sample tile buffers in a ring (tile q of the CTA lives in buffer q % NB).
//   * work item = (tile q, slice p): 32 of the tile's 160 (sample, channel) columns.  Math warps pull items in order from a
//     shared-memory counter, so a warp that runs ahead (the scheduler favours high warp ids; SMSP 3 hosts one math warp
//     less) simply takes more items -- nobody waits at a barrier for the slowest warp of a group.
//   * a math warp waits for full[q % NB] (TMA bytes landed + trig table written), runs the backward chain on its 32
//     columns (spectrum gradient in place, the column's generator forms T_x, T_y, T_z to gp), and arrives on empty[q % NB]
//     (count 5).
//   * the producer warps (the last warps of the CTA: highest scheduling priority) own everything else.  For tile r,
//     in ring order: wait empty -> column sums of the finished tile into 3 float2 registers per lane (this IS the batch
//     reduction of the item_rep gradient: no atomics, no L2 reduce traffic, fixed order -> bit-reproducible) -> TMA bulk load of tile r + NB into the buffer (plus an
//     L2 prefetch of the tile after it) -> write the trig table of tile r + NB, which they computed *before* the wait
//     from angles fetched one tile earlier still -> arrive on full; in the load's shadow: T_k summed over the 10 channels,
//     the body-frame relation, g_angles stored.  The buffer's turnaround is column sums + one load latency; the other NB - 1
//     tiles are being computed or waiting meanwhile.
// Requires full 16-sample tiles and 16-byte aligned g_y (the host sends a ragged tail through wigner_bwd_kernel).
// 11 math + 5 producer warps = 16 warps x 128 registers: the whole register file (warps are allocated in fours).  The split is
// measured: since the angle gradients stopped needing the forward recompute the buffer turnaround (column sums + refill) binds,
// and 13+3 / 12+4 / 11+5 / 10+6 / 9+7 warps run a 2^18-sample launch in 0.219 / 0.210 / 0.199 / 0.201 / 0.205 ms.
#ifndef WD_MATH_WARPS_N
#define WD_MATH_WARPS_N 11
#endif
#ifndef WD_TOTAL_WARPS_N
#define WD_TOTAL_WARPS_N 16
#endif
constexpr int WD_S = 16, WD_MATH_WARPS = WD_MATH_WARPS_N, WD_PROD_WARPS = WD_TOTAL_WARPS_N - WD_MATH_WARPS, WD_THREADS = WD_TOTAL_WARPS_N * 32, WD_SLICES = 5;
constexpr int WD_PL = WD_PROD_WARPS * 32;       // producer lanes
// tile buffers in the ring: as many as shared memory holds (degrees 0..8: 4 x 51.8 KB, degrees 0..6: 6 x 31.4 KB)
__host__ __device__ constexpr int wd_bufs(int LT) { return LT <= 6 ? 6 : 4; }


template <int CT, int LT>
__global__ void __launch_bounds__(WD_THREADS, 1)
wigner_bwd_ws_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
                     float* __restrict__ gangles, float* __restrict__ partial, int64_t ntiles, int transpose) {
    constexpr int C = CT, M = (LT + 1) * (LT + 1), MC = M * C, COLS = WD_S * C;     // COLS = 160 columns per tile
    constexpr int WD_BUFS = wd_bufs(LT);
    static_assert(COLS == WD_SLICES * 32, "a tile must split into whole warps");
    static_assert(MC % 2 == 0, "column sums read float2");
    constexpr uint32_t TILE_BYTES = WD_S * MC * 4u;
    constexpr int NC2 = MC / 2, KACC = (NC2 + WD_PL - 1) / WD_PL;                             // float2 columns, accumulators per lane
    extern __shared__ __align__(16) float smem[];
    float* tiles = smem;                                             // [4][16][MC]
    float* trig_all = tiles + WD_BUFS * WD_S * MC;                   // [4][16][52]
    float* gp_all = trig_all + WD_BUFS * WD_S * WG_TRIG_STRIDE;      // [4][160][3]
    uint64_t* full = reinterpret_cast<uint64_t*>(gp_all + WD_BUFS * COLS * 3);      // [4]  count 3: expect_tx arrive + one trig arrive per producer warp
    uint64_t* empty = full + WD_BUFS;                                               // [4]  count 5: one per slice
    int* s_next = reinterpret_cast<int*>(empty + WD_BUFS);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int b = 0; b < WD_BUFS; ++b) { mbar_init(full + b, 1 + WD_PROD_WARPS); mbar_init(empty + b, WD_SLICES); }
        *s_next = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < WD_MATH_WARPS) {
        // ------------------------------------------------------------ math warps
        for (;;) {
            int item = 0;
            if (lane == 0) item = atomicAdd(s_next, 1);
            item = __shfl_sync(0xffffffffu, item, 0);
            const int64_t q = item / WD_SLICES;
            if (q >= my_tiles) break;
            const int p = item - int(q) * WD_SLICES;
            const int buf = int(q % WD_BUFS);
            const int t = p * 32 + lane, s = t / C, c = t - s * C;
            mbar_wait(full + buf, uint32_t(q / WD_BUFS) & 1u);
            GenAcc acc;
            bwd_gen_unrolled<0, LT, true>(spectrum + c, tiles + buf * WD_S * MC + s * MC + c, C,
                                          reinterpret_cast<const float4*>(trig_all + (buf * WD_S + s) * WG_TRIG_STRIDE), acc);
            float* gp = gp_all + (buf * COLS + t) * 3;          // (T_x, T_y, T_z) of this column
            gp[0] = acc.tx;
            gp[1] = acc.ty;
            gp[2] = acc.tz();
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + buf);       // release: the slice's tile and gp writes are visible to the producer
        }
    } else {
        // ------------------------------------------------------------ producer warps (pw = 0, 1)
        const int pw = warp - WD_MATH_WARPS, pl = pw * 32 + lane;       // lane of the producer group
        f32x2_t acc[KACC];
#pragma unroll
        for (int k = 0; k < KACC; ++k) acc[k] = 0ull;
        // (sample, angle) pair pl (< 48) of a tile is this lane's trig job.  The angle of tile j + 1 is fetched while the
        // table of tile j is computed, and the table is computed before the buffer is awaited: neither the DRAM latency
        // nor sincosf sits between a buffer becoming free and its refill.
        const bool trig_lane = pl < WD_S * 3;
        const int t_ss = pl / 3, t_a = pl - 3 * t_ss;
        auto load_phi = [&](int64_t j) -> float {      // angle of this lane's pair in the CTA's tile j (0 past the end)
            if (!trig_lane || j >= my_tiles) return 0.f;
            const int64_t n0 = (first + j * stride) * WD_S;
            return transpose ? -__ldg(angles + (n0 + t_ss) * 3 + (2 - t_a)) : __ldg(angles + n0 * 3 + pl);
        };
        float tr[WG_TRIG_ANGLE];
        auto trig_store = [&](int buf) {
            if (trig_lane) {
                float4* d = reinterpret_cast<float4*>(trig_all + (buf * WD_S + t_ss) * WG_TRIG_STRIDE + t_a * WG_TRIG_ANGLE);
#pragma unroll
                for (int k = 0; k < 4; ++k) d[k] = make_float4(tr[4 * k], tr[4 * k + 1], tr[4 * k + 2], tr[4 * k + 3]);
            }
        };
        auto issue_load = [&](int64_t j) {      // one lane: first arrival on full + the bulk load; L2 prefetch of the tile after it
            const int buf = int(j % WD_BUFS);
            mbar_expect_tx(full + buf, TILE_BYTES);
            tma_load(tiles + buf * WD_S * MC, gout + (first + j * stride) * WD_S * MC, TILE_BYTES, full + buf);
            if (j + 1 < my_tiles)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                             :: "l"(gout + (first + (j + 1) * stride) * WD_S * MC), "r"(TILE_BYTES) : "memory");
        };
        // prologue: fill the ring
        float phi = load_phi(0);
        for (int64_t j = 0; j < WD_BUFS && j < my_tiles; ++j) {
            if (pl == 0) issue_load(j);
            const float phi_next = load_phi(j + 1);
            trig_fill(tr, phi);
            trig_store(int(j));
            __syncwarp();
            if (lane == 0) mbar_arrive(full + int(j));
            phi = phi_next;
        }
        // phi now belongs to tile min(WD_BUFS, my_tiles)
        for (int64_t r = 0; r < my_tiles; ++r) {
            const int buf = int(r % WD_BUFS);
            const int64_t jn = r + WD_BUFS;
            if (jn < my_tiles) {
                const float phi_next = load_phi(jn + 1);
                trig_fill(tr, phi);
                phi = phi_next;
            }
            mbar_wait(empty + buf, uint32_t(r / WD_BUFS) & 1u);
            // batch reduction: column sums of the finished tile, float2 columns pl, pl + 64, ... (rows are 8-byte aligned)
            const float* tile = tiles + buf * WD_S * MC;
            {
#pragma unroll
                for (int row = 0; row < WD_S; ++row) {
#pragma unroll
                    for (int k = 0; k < KACC; ++k) {
                        const int col2 = pl + WD_PL * k;
                        if (col2 < NC2) {
                            const float2 v = *reinterpret_cast<const float2*>(tile + row * MC + 2 * col2);
                            asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[k]) : "l"(wg2::pk(v.x, v.y)));
                        }
                    }
                }
            }
            // the buffer is free as soon as every producer warp has summed its columns: refill it first, the angle gradients
            // (which read only gp and the trig table) follow off the load's critical path
            if (jn < my_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of the buffer before the bulk write
                named_bar_sync(1, WD_PL);                                      // all producer warps are done with the buffer
                if (pl == 0) issue_load(jn);
            }
            // angle gradients of sample t_ss: T_k summed over the channels, then the body-frame relation for the effective angles
            // (a', b', c') whose trig table is still in the buffer; transpose maps (a', b', c') = (-c, -b, -a) back
            if (trig_lane) {
                const int64_t n0 = (first + r * stride) * WD_S;
                const float* gp = gp_all + buf * COLS * 3 + t_ss * C * 3;
                float tx = 0.f, ty = 0.f, tz = 0.f;
#pragma unroll
                for (int cc = 0; cc < C; ++cc) { tx += gp[cc * 3]; ty += gp[cc * 3 + 1]; tz += gp[cc * 3 + 2]; }
                const float* tr_s = trig_all + (buf * WD_S + t_ss) * WG_TRIG_STRIDE;
                const float gval = angle_grad_from_generators(transpose ? 2 - t_a : t_a, tx, ty, tz, tr_s[WG_TRIG_ANGLE], tr_s[WG_TRIG_ANGLE + 2],
                                                              tr_s[2 * WG_TRIG_ANGLE], tr_s[2 * WG_TRIG_ANGLE + 2]);
                gangles[n0 * 3 + pl] = transpose ? -gval : gval;
            }
            if (jn < my_tiles) {
                if (pw < 2) named_bar_sync(2, 64);     // the 48 trig lanes sit in producer warps 0 and 1: old table read before it is replaced
                trig_store(buf);
                __syncwarp();
                if (lane == 0) mbar_arrive(full + buf);
            }
        }
        float* prow = partial + int64_t(blockIdx.x) * MC;
#pragma unroll
        for (int k = 0; k < KACC; ++k) {
            const int col2 = pl + WD_PL * k;
            if (col2 < NC2) *reinterpret_cast<float2*>(prow + 2 * col2) = make_float2(wg2::plo(acc[k]), wg2::phi(acc[k]));
        }
    }
}

// partial [nblk][MC] -> out[MC] (accumulate: out += ...); block = (32, 8): 32 consecutive columns, 8-way split over rows.
__global__ void __launch_bounds__(256)
wigner_reduce_partials(const float* __restrict__ partial, float* __restrict__ out, int nblk, int MC, int accumulate) {
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
        out[o] = accumulate ? out[o] + a : a;
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

struct WgGeom { int M, MC, S, threads, sms; size_t smem_fwd, smem_bwd; int64_t ntiles; };

// S = samples per CTA: about 192 threads (160 for C = 10), tile <= ~52 KB so that several CTAs share an SM.
static int wigner_geometry(const char* name, int64_t N, int lmin, int lmax, int C, bool shared, WgGeom& g) {
    if (N < 0 || C <= 0 || lmin < 0 || lmax < lmin) { set_error("%s: bad sizes (N=%lld, C=%d, degrees %d..%d)", name, (long long)N, C, lmin, lmax); return LV_ERR_ARG; }
    if (lmax > WG_LMAX) { set_error("%s: degree %d > %d is not supported by the unrolled kernels", name, lmax, WG_LMAX); return LV_ERR_UNSUPPORTED; }
    if (C > WG_MAX_THREADS) { set_error("%s: more than %d channels unsupported", name, WG_MAX_THREADS); return LV_ERR_UNSUPPORTED; }
    DevInfo di = dev_info();
    if (!di.ok) { set_error("%s: cannot query the CUDA device", name); return int(cudaErrorInvalidDevice); }
    g.sms = di.sms;
    g.M = (lmax + 1) * (lmax + 1) - lmin * lmin;
    g.MC = g.M * C;
    int S = (C == 10 ? wg_threads(10) : 192) / C;    // the C = 10 specialisations are bounded to 160 threads
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
    return LV_OK;
}

template <typename K>
static int opt_in_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return LV_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes));
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%zu): %s", bytes, cudaGetErrorString(e)); return int(e); }
    return LV_OK;
}

template <bool SHARED, int CT, int LT>
static int launch_fwd(const WgGeom& g, const float* angles, const float* spectrum, float* out, int64_t N, int lmin,
                      int lmax, int C, int transpose, cudaStream_t st) {
    int rc = opt_in_smem(wigner_fwd_kernel<SHARED, CT, LT>, g.smem_fwd);
    if (rc) return rc;
    wigner_fwd_kernel<SHARED, CT, LT><<<unsigned(g.ntiles), g.threads, g.smem_fwd, st>>>(angles, spectrum, out, N, lmin, lmax, C, g.S, transpose);
    return check_launch("wigner_apply_fwd");
}

// persistent grid = exactly the CTAs that are resident at once (a second partial wave would run alone)
template <bool SHARED, int CT, int LT>
static int launch_bwd(const WgGeom& g, const float* angles, const float* spectrum, const float* gout, float* gangles,
                      float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int lmin, int lmax, int C,
                      int transpose, cudaStream_t st, int* grid_out) {
    int rc = opt_in_smem(wigner_bwd_kernel<SHARED, CT, LT>, g.smem_bwd);
    if (rc) return rc;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wigner_bwd_kernel<SHARED, CT, LT>, g.threads, g.smem_bwd);
    if (e != cudaSuccess || per_sm < 1) { set_error("wigner_apply_bwd: occupancy query failed (%s)", cudaGetErrorString(e)); return e != cudaSuccess ? int(e) : LV_ERR_UNSUPPORTED; }
    if (per_sm > WG_MAX_CTAS_PER_SM) per_sm = WG_MAX_CTAS_PER_SM;
    const int64_t cap = int64_t(g.sms) * per_sm;
    const int grid = int(g.ntiles < cap ? g.ntiles : cap);
    if (SHARED && (!workspace || workspace_floats < int64_t(grid) * g.MC)) {
        set_error("wigner_apply_bwd: workspace of %lld floats required", (long long)(int64_t(grid) * g.MC));
        return LV_ERR_ARG;
    }
    *grid_out = grid;
    wigner_bwd_kernel<SHARED, CT, LT><<<grid, g.threads, g.smem_bwd, st>>>(angles, spectrum, gout, gangles, SHARED ? nullptr : gspectrum,
                                                                           SHARED ? workspace : nullptr, N, lmin, lmax, C, g.S, transpose, g.ntiles);
    return check_launch("wigner_apply_bwd");
}

// ---- warp-decoupled backward (shared spectrum, C = 10, degrees 0..8 or 0..6, 16-byte aligned g_y) -----------------

static bool ws_bwd_eligible(int C, int lmin, int lmax, const float* gout, int64_t N) {
    return C == 10 && lmin == 0 && (lmax == 8 || lmax == 6) && N >= WD_S && (reinterpret_cast<uintptr_t>(gout) & 15u) == 0;
}
// workspace rows (of MC floats): one partial row per CTA [sms] | tail partial [1]
static int64_t ws_bwd_workspace_rows(int sms) { return int64_t(sms) + 1; }

template <int LT>
static int launch_bwd_ws(const WgGeom& g, const float* angles, const float* spectrum, const float* gout, float* gangles,
                         float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int transpose, int accumulate,
                         cudaStream_t st) {
    constexpr int C = 10, MC = (LT + 1) * (LT + 1) * C;
    if (!workspace || workspace_floats < ws_bwd_workspace_rows(g.sms) * MC) {
        set_error("wigner_apply_bwd: workspace of %lld floats required", (long long)(ws_bwd_workspace_rows(g.sms) * MC));
        return LV_ERR_ARG;
    }
    const int64_t ntiles = N / WD_S, n_full = ntiles * WD_S, n_tail = N - n_full;
    const int grid = int(ntiles < g.sms ? ntiles : g.sms);
    constexpr int NB = wd_bufs(LT);
    const size_t smem = size_t(NB * WD_S * MC + NB * WD_S * WG_TRIG_STRIDE + NB * WD_S * C * 3) * 4 + 2 * NB * 8 + 16;
    int rc = opt_in_smem(wigner_bwd_ws_kernel<C, LT>, smem);
    if (rc) return rc;
    float* partial = workspace;      // one row per CTA, then the tail's rows
    wigner_bwd_ws_kernel<C, LT><<<grid, WD_THREADS, smem, st>>>(angles, spectrum, gout, gangles, partial, ntiles, transpose);
    if ((rc = check_launch("wigner_apply_bwd (ws)"))) return rc;
    int rows = grid;
    if (n_tail > 0) {
        WgGeom gt = g;
        gt.ntiles = (n_tail + g.S - 1) / g.S;
        int tail_grid = 0;
        rc = launch_bwd<true, 0, -1>(gt, angles + n_full * 3, spectrum, gout + n_full * MC, gangles + n_full * 3, nullptr,
                                     partial + int64_t(rows) * MC, int64_t(gt.ntiles) * MC, n_tail, 0, LT, C, transpose, st, &tail_grid);
        if (rc) return rc;
        rows += tail_grid;
    }
    wigner_reduce_partials<<<(MC + 31) / 32, dim3(32, 8), 0, st>>>(partial, gspectrum, rows, MC, accumulate);
    return check_launch("wigner_reduce_partials");
}

#include "wigner_bwd_dg.cuh"

// ---- degree-specialised backward (shared spectrum, C = 10, degrees 0..8 or 0..6, 16-byte aligned g_y, N >= 2) ------
// LV_WIGNER_BWD=ws in the environment selects the first TMA-fed kernel instead (A/B measurements on one box).
static bool dg_bwd_enabled() {
    static const int on = [] { const char* e = getenv("LV_WIGNER_BWD"); return (e && e[0] == 'w') ? 0 : 1; }();
    return on != 0;
}
static bool dg_bwd_eligible(int C, int lmin, int lmax, const float* gout, int64_t N) {
    return dg_bwd_enabled() && C == 10 && lmin == 0 && (lmax == 8 || lmax == 6) && N >= 2 && (reinterpret_cast<uintptr_t>(gout) & 15u) == 0;
}

template <class CFG>
static int launch_bwd_dg(const WgGeom& g, const float* angles, const float* spectrum, const float* gout, float* gangles,
                         float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int transpose, int accumulate,
                         cudaStream_t st) {
    using G = dg::Geo<CFG>;
    constexpr int MC = G::MC, S = CFG::S;
    if (!workspace || workspace_floats < ws_bwd_workspace_rows(g.sms) * MC) {
        set_error("wigner_apply_bwd: workspace of %lld floats required", (long long)(ws_bwd_workspace_rows(g.sms) * MC));
        return LV_ERR_ARG;
    }
    // bulk copies move multiples of 16 bytes = an even number of 4*MC-byte rows: an odd last sample goes to the generic kernel
    const int64_t n_main = N & ~int64_t(1), n_tail = N - n_main;
    const int64_t ntiles = (n_main + S - 1) / S;
    const int last_rows = int(n_main - (ntiles - 1) * S);
    const int grid = int(ntiles < g.sms ? ntiles : g.sms);
    int rc = opt_in_smem(dg::wigner_bwd_dg_kernel<CFG>, G::SMEM);
    if (rc) return rc;
    float* partial = workspace;      // one row per CTA, then the tail's row
    dg::wigner_bwd_dg_kernel<CFG><<<grid, G::THREADS, G::SMEM, st>>>(angles, spectrum, gout, gangles, partial, ntiles, last_rows, transpose);
    if ((rc = check_launch("wigner_apply_bwd (dg)"))) return rc;
    int rows = grid;
    if (n_tail > 0) {
        WgGeom gt = g;
        gt.ntiles = (n_tail + g.S - 1) / g.S;
        int tail_grid = 0;
        rc = launch_bwd<true, 0, -1>(gt, angles + n_main * 3, spectrum, gout + n_main * MC, gangles + n_main * 3, nullptr,
                                     partial + int64_t(rows) * MC, int64_t(gt.ntiles) * MC, n_tail, 0, CFG::LT, 10, transpose, st, &tail_grid);
        if (rc) return rc;
        rows += tail_grid;
    }
    wigner_reduce_partials<<<(MC + 31) / 32, dim3(32, 8), 0, st>>>(partial, gspectrum, rows, MC, accumulate);
    return check_launch("wigner_reduce_partials");
}

}  // namespace lv

// compile-time specialisations: 10 channels (ActionNet default, decoders.py:11 / main.py:168) with degrees
// 0..8 (BASELINE configs 3, 5) or 0..6 (config 4); everything else takes the run-time instantiation.
#define WG_DISPATCH(C, lmin, lmax, SHARED, FN, ...)                              \
    ((C) == 10 && (lmin) == 0 && (lmax) == 8 ? FN<SHARED, 10, 8>(__VA_ARGS__)    \
     : (C) == 10 && (lmin) == 0 && (lmax) == 6 ? FN<SHARED, 10, 6>(__VA_ARGS__)  \
                                               : FN<SHARED, 0, -1>(__VA_ARGS__))

// ====================================================================== C ABI
extern "C" int64_t lv_wigner_bwd_workspace_floats(int64_t N, int lmin, int lmax, int C) {
    lv::WgGeom g;
    if (lv::wigner_geometry("wigner_bwd_workspace", N, lmin, lmax, C, true, g) != LV_OK) return -1;
    const int64_t cap = int64_t(g.sms) * lv::WG_MAX_CTAS_PER_SM;
    int64_t rows = g.ntiles < cap ? g.ntiles : cap;
    if (C == 10 && lmin == 0 && (lmax == 8 || lmax == 6) && N >= 2) {
        const int64_t tma_rows = lv::ws_bwd_workspace_rows(g.sms);
        if (tma_rows > rows) rows = tma_rows;
    }
    return rows * g.MC;
}

extern "C" int lv_wigner_apply_fwd_f32(const float* angles, const float* spectrum, float* out, int64_t N, int lmin,
                                       int lmax, int C, int shared_spectrum, int transpose, void* stream) {
    lv::WgGeom g;
    int rc = lv::wigner_geometry("wigner_apply_fwd", N, lmin, lmax, C, shared_spectrum != 0, g);
    if (rc) return rc;
    if (N == 0) return LV_OK;
    if (!angles || !spectrum || !out) { lv::set_error("wigner_apply_fwd: null pointer"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (shared_spectrum) return WG_DISPATCH(C, lmin, lmax, true, lv::launch_fwd, g, angles, spectrum, out, N, lmin, lmax, C, transpose, st);
    return WG_DISPATCH(C, lmin, lmax, false, lv::launch_fwd, g, angles, spectrum, out, N, lmin, lmax, C, transpose, st);
}

extern "C" int lv_wigner_recon_sse_f32(const float* angles, const float* item_rep, const float* x, float* out, int64_t N, int64_t B,
                                       int lmin, int lmax, int C, int transpose, void* stream) {
    lv::WgGeom g;
    int rc = lv::wigner_geometry("wigner_recon_sse", N, lmin, lmax, C, true, g);
    if (rc) return rc;
    if (N == 0) return LV_OK;
    if (B <= 0 || !angles || !item_rep || !x || !out) { lv::set_error("wigner_recon_sse: null pointer or empty batch"); return LV_ERR_ARG; }
    const size_t smem = size_t(lv::align4i(g.S * g.MC) + g.S * lv::WG_TRIG_STRIDE + g.S * C) * 4;
    rc = lv::opt_in_smem(lv::wigner_sse_kernel, smem);
    if (rc) return rc;
    lv::wigner_sse_kernel<<<unsigned(g.ntiles), g.threads, smem, reinterpret_cast<cudaStream_t>(stream)>>>(angles, item_rep, x, out, N, B, lmin, lmax, C, g.S, transpose);
    return lv::check_launch("wigner_recon_sse");
}

extern "C" int lv_wigner_apply_bwd_f32(const float* angles, const float* spectrum, const float* gout, float* gangles,
                                       float* gspectrum, float* workspace, int64_t workspace_floats, int64_t N, int lmin,
                                       int lmax, int C, int shared_spectrum, int transpose, void* stream) {
    lv::WgGeom g;
    int rc = lv::wigner_geometry("wigner_apply_bwd", N, lmin, lmax, C, shared_spectrum != 0, g);
    if (rc) return rc;
    if (!gspectrum) { lv::set_error("wigner_apply_bwd: null pointer"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int accumulate = (shared_spectrum & 2) ? 1 : 0;     // bit 1: gspectrum += batch sum (micro-batched steps)
    if (accumulate && !(shared_spectrum & 1)) { lv::set_error("wigner_apply_bwd: accumulate needs a shared spectrum"); return LV_ERR_ARG; }
    if (N == 0) {
        if (shared_spectrum && !accumulate) {
            cudaError_t e = cudaMemsetAsync(gspectrum, 0, size_t(g.MC) * 4, st);
            if (e != cudaSuccess) { lv::set_error("wigner_apply_bwd: memset: %s", cudaGetErrorString(e)); return int(e); }
        }
        return LV_OK;
    }
    if (!angles || !spectrum || !gout || !gangles) { lv::set_error("wigner_apply_bwd: null pointer"); return LV_ERR_ARG; }
    int grid = 0;
    if (shared_spectrum && lv::dg_bwd_eligible(C, lmin, lmax, gout, N)) {
        if (lmax == 8) return lv::launch_bwd_dg<lv::dg::LV_DG_CFG8>(g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, transpose, accumulate, st);
        return lv::launch_bwd_dg<lv::dg::LV_DG_CFG6>(g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, transpose, accumulate, st);
    }
    if (shared_spectrum && lv::ws_bwd_eligible(C, lmin, lmax, gout, N)) {
        if (lmax == 8) return lv::launch_bwd_ws<8>(g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, transpose, accumulate, st);
        return lv::launch_bwd_ws<6>(g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, transpose, accumulate, st);
    }
    if (shared_spectrum) {
        rc = WG_DISPATCH(C, lmin, lmax, true, lv::launch_bwd, g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, lmin, lmax, C, transpose, st, &grid);
        if (rc) return rc;
        lv::wigner_reduce_partials<<<(g.MC + 31) / 32, dim3(32, 8), 0, st>>>(workspace, gspectrum, grid, g.MC, accumulate);
        return lv::check_launch("wigner_reduce_partials");
    }
    return WG_DISPATCH(C, lmin, lmax, false, lv::launch_bwd, g, angles, spectrum, gout, gangles, gspectrum, workspace, workspace_floats, N, lmin, lmax, C, transpose, st, &grid);
}
