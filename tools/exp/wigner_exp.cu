// Experimental Wigner forward variants for ablation timing (NOT product code; built by tools/exp/run_exp.py).
#include "../../lie_vae_b200/csrc/common.cuh"
#include "wigner_gen_scalar.cuh"
#include <cstdio>

namespace lv {
void set_error(const char*, ...) {}
int check_launch(const char*) { return int(cudaGetLastError()); }
}
using namespace lv;
using lv::wg::jmul;

constexpr int LMAX = 8, TS = 52, C = 10, M = 81, MC = 810;

template <int L, bool REG>
__device__ __forceinline__ void xrot(float (&x)[2 * L + 1], const float2* __restrict__ cs) {
    if constexpr (REG) {
#pragma unroll
        for (int m = 1; m <= L; ++m) {
            const float2 t = cs[m - 1];
            const float a = x[L - m], b = x[L + m];
            x[L - m] = fmaf(t.x, a, t.y * b);
            x[L + m] = fmaf(t.x, b, -(t.y * a));
        }
    } else {
        const float4* cs4 = reinterpret_cast<const float4*>(cs);
#pragma unroll
        for (int p = 0; p < (L + 1) / 2; ++p) {
            float4 t = cs4[p];
            { const int m = 2 * p + 1; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.x, a, t.y * b); x[L + m] = fmaf(t.x, b, -(t.y * a)); }
            if (2 * p + 2 <= L) { const int m = 2 * p + 2; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.z, a, t.w * b); x[L + m] = fmaf(t.z, b, -(t.w * a)); }
        }
    }
}

__device__ __forceinline__ void stage_trig(float* s_trig, const float* angles, int64_t n0, int rows) {
    for (int j = threadIdx.x; j < rows * 3; j += blockDim.x) {
        const int s = j / 3, a = j - 3 * s;
        float s1, c1;
        sincosf(__ldg(angles + n0 * 3 + j), &s1, &c1);
        float2* dst = reinterpret_cast<float2*>(s_trig + s * TS + a * 16);
        float cm = c1, sm = s1;
#pragma unroll
        for (int m = 1; m <= LMAX; ++m) { dst[m - 1] = make_float2(cm, sm); const float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn; }
    }
}

// VAR bits: 1 = no copy-out, 2 = trig in registers, 4 = item from smem, 8 = no item loads (constant), 16 = no trig loads (constant)
template <int L, int VAR>
__device__ __forceinline__ void degree_fwd(const float* src, float* dst, const float2* tg) {
    float x[2 * L + 1], y[2 * L + 1];
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = (VAR & 8) ? 0.5f + i : ((VAR & 4) ? src[i * C] : __ldg(src + i * C));
    constexpr bool REG = (VAR & 2) || (VAR & 16);
    xrot<L, REG>(x, tg + 16);
    jmul<L>(x, y);
    xrot<L, REG>(y, tg + 8);
    jmul<L>(y, x);
    xrot<L, REG>(x, tg);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) dst[i * C] = x[i];
}

template <int L, int VAR>
__device__ __forceinline__ void degrees_from(const float* srow, float* trow, const float2* tg) {
    degree_fwd<L, VAR>(srow + L * L * C, trow + L * L * C, tg);
    if constexpr (L < LMAX) degrees_from<L + 1, VAR>(srow, trow, tg);
}

template <int VAR, int MINB>
__global__ void __launch_bounds__(160, MINB)
fwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, float* __restrict__ out, int64_t N, int S) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float* s_trig = smem + S * MC;
    float* s_item = s_trig + S * TS;
    const int64_t n0 = int64_t(blockIdx.x) * S;
    const int rows = int(min(int64_t(S), N - n0));
    stage_trig(s_trig, angles, n0, rows);
    if (VAR & 4) for (int o = threadIdx.x; o < MC; o += blockDim.x) s_item[o] = __ldg(spectrum + o);
    __syncthreads();
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    if (s < rows) {
        const float2* tg = reinterpret_cast<const float2*>(s_trig + s * TS);
        float2 tr[24];
        if (VAR & 2) {
            const float4* t4 = reinterpret_cast<const float4*>(tg);
#pragma unroll
            for (int i = 0; i < 12; ++i) { const float4 v = t4[i]; tr[2 * i] = make_float2(v.x, v.y); tr[2 * i + 1] = make_float2(v.z, v.w); }
        } else if (VAR & 16) {
#pragma unroll
            for (int i = 0; i < 24; ++i) tr[i] = make_float2(0.8f + 0.001f * (t + i), 0.6f - 0.001f * i);
        }
        const float2* tp = ((VAR & 2) || (VAR & 16)) ? tr : tg;
        float* trow = tile + s * MC + c;
        const float* srow = (VAR & 4) ? s_item + c : spectrum + c;
        degrees_from<0, VAR>(srow, trow, tp);
    }
    if (VAR & 32) {
        // TMA bulk store of the whole tile: writers fence generic->async proxy, one thread issues the copy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t bytes = uint32_t(rows) * MC * 4u;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + n0 * MC), "r"(uint32_t(__cvta_generic_to_shared(tile))), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        return;
    }
    __syncthreads();
    if (!(VAR & 1)) tile_s2g(out + n0 * MC, tile, rows * MC);
    else if (threadIdx.x == 0) out[n0 * MC] = tile[threadIdx.x];
}

template <int VAR, int MINB>
static int launch(const float* angles, const float* spectrum, float* out, int64_t N, int S, cudaStream_t st) {
    const size_t smem = size_t(S * MC + S * TS + MC) * 4;
    cudaFuncSetAttribute(fwd_kernel<VAR, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    fwd_kernel<VAR, MINB><<<unsigned((N + S - 1) / S), S * C, smem, st>>>(angles, spectrum, out, N, S);
    return int(cudaGetLastError());
}

extern "C" int exp_wigner_fwd(int var, const float* angles, const float* spectrum, float* out, int64_t N, int S, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (var) {
        case 0: return launch<0, 1>(angles, spectrum, out, N, S, st);
        case 1: return launch<1, 1>(angles, spectrum, out, N, S, st);
        case 2: return launch<2, 3>(angles, spectrum, out, N, S, st);
        case 4: return launch<4, 1>(angles, spectrum, out, N, S, st);
        case 6: return launch<6, 3>(angles, spectrum, out, N, S, st);
        case 8: return launch<8, 1>(angles, spectrum, out, N, S, st);
        case 16: return launch<16, 3>(angles, spectrum, out, N, S, st);
        case 24: return launch<24, 3>(angles, spectrum, out, N, S, st);
        case 25: return launch<25, 3>(angles, spectrum, out, N, S, st);
        case 32: return launch<32, 1>(angles, spectrum, out, N, S, st);
        case 33: return launch<32, 4>(angles, spectrum, out, N, S, st);
        default: return -1;
    }
}

// ===================================================================== backward variants
template <int L>
__device__ __forceinline__ float gdot(const float (&h)[2 * L + 1], const float (&w)[2 * L + 1]) {
    float acc = 0.f;
#pragma unroll
    for (int m = 1; m <= L; ++m) acc = fmaf(float(m), fmaf(h[L - m], w[L + m], -(h[L + m] * w[L - m])), acc);
    return acc;
}
template <int L>
__device__ __forceinline__ void xrot_t(float (&x)[2 * L + 1], const float2* __restrict__ cs) {
    const float4* cs4 = reinterpret_cast<const float4*>(cs);
#pragma unroll
    for (int p = 0; p < (L + 1) / 2; ++p) {
        float4 t = cs4[p];
        { const int m = 2 * p + 1; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.x, a, -(t.y * b)); x[L + m] = fmaf(t.x, b, t.y * a); }
        if (2 * p + 2 <= L) { const int m = 2 * p + 2; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.z, a, -(t.w * b)); x[L + m] = fmaf(t.z, b, t.w * a); }
    }
}

// BV bits: 1 = skip tile load, 2 = skip reduce phase, 4 = skip math (copy g through), 8 = item via LDG instead of smem
template <int L, int BV>
__device__ __forceinline__ void degree_bwd(const float* src, float* g, const float2* tg, float& ga, float& gb, float& gc) {
    float x[2 * L + 1], y[2 * L + 1], w2[2 * L + 1];
    if (BV & 4) {
#pragma unroll
        for (int i = 0; i < 2 * L + 1; ++i) g[i * C] = g[i * C] * 1.0001f;
        return;
    }
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = (BV & 8) ? __ldg(src + i * C) : src[i * C];
    xrot<L, false>(x, tg + 16);
    jmul<L>(x, w2);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = w2[i];
    xrot<L, false>(y, tg + 8);
    jmul<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = g[i * C];
    xrot_t<L>(y, tg);
    ga += gdot<L>(y, x);
    jmul<L>(y, x);
    xrot_t<L>(x, tg + 8);
    gb += gdot<L>(x, w2);
    jmul<L>(x, y);
    xrot_t<L>(y, tg + 16);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = (BV & 8) ? __ldg(src + i * C) : src[i * C];
    gc += gdot<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) g[i * C] = y[i];
}
template <int L, int BV>
__device__ __forceinline__ void degrees_bwd_from(const float* srow, float* trow, const float2* tg, float& ga, float& gb, float& gc) {
    degree_bwd<L, BV>(srow + L * L * C, trow + L * L * C, tg, ga, gb, gc);
    if constexpr (L < LMAX) degrees_bwd_from<L + 1, BV>(srow, trow, tg, ga, gb, gc);
}

template <int BV, int MINB>
__global__ void __launch_bounds__(160, MINB)
bwd_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
           float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int S, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float* s_trig = tile + S * MC;
    float* s_gp = s_trig + S * TS;
    float* s_acc = s_gp + S * C * 3 + 4;
    float* s_item = s_acc + MC + 2;
    const int t = threadIdx.x;
    const int s = t / C, c = t - s * C;
    for (int o = t; o < MC; o += blockDim.x) { s_acc[o] = 0.f; s_item[o] = __ldg(spectrum + o); }
    if (BV & 64) {
        // de-synchronise the CTAs that share an SM: group g = blockIdx / #SMs starts g * T/3 later
        const long long wait = (long long)(blockIdx.x / 148) * 6000;
        const long long t0 = clock64();
        while (clock64() - t0 < wait) { }
    }
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
        const int64_t n0 = tile_idx * S;
        const int rows = int(min(int64_t(S), N - n0));
        if (BV & 16) {
            tile_g2s(tile, gout + n0 * MC, rows * MC);
            const int64_t nxt = tile_idx + gridDim.x;
            if (t == 0 && nxt < ntiles) {
                const int nrows = int(min(int64_t(S), N - nxt * S));
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(gout + nxt * S * MC), "r"(uint32_t(nrows) * MC * 4u) : "memory");
            }
            stage_trig(s_trig, angles, n0, rows);
            tile_async_wait();
        } else {
            stage_trig(s_trig, angles, n0, rows);
            if (!(BV & 1)) { tile_g2s(tile, gout + n0 * MC, rows * MC); tile_async_wait(); }
        }
        __syncthreads();
        if (s < rows) {
            const float2* tg = reinterpret_cast<const float2*>(s_trig + s * TS);
            float* trow = tile + s * MC + c;
            const float* srow = (BV & 8) ? spectrum + c : s_item + c;
            float ga = 0.f, gb = 0.f, gc = 0.f;
            degrees_bwd_from<0, BV>(srow, trow, tg, ga, gb, gc);
            s_gp[t * 3 + 0] = ga; s_gp[t * 3 + 1] = gb; s_gp[t * 3 + 2] = gc;
        }
        if (BV & 128) {
            // TMA bulk reduce-add: the copy engine adds the whole tile into this CTA's fp32 accumulator in global
            // memory (L2); the SM only fences, issues one instruction and waits for the smem read.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (t == 0) {
                float* gacc = partial + int64_t(blockIdx.x) * S * MC;
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             :: "l"(gacc), "r"(uint32_t(__cvta_generic_to_shared(tile))), "r"(uint32_t(rows) * MC * 4u) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        } else {
        __syncthreads();
        }
        if (BV & 128) {
        } else if (BV & 32) {
            // 128-bit column sums: 202 float4 columns + one float2 tail (MC = 810)
            for (int q = t; q < 203; q += blockDim.x) {
                if (q < 202) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int r = 0; r < rows; ++r) {
                        // rows start at r*3240 B: 16B-aligned only for even r -> use two float2 loads
                        const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q);
                        const float2 v = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q + 2);
                        a.x += u.x; a.y += u.y; a.z += v.x; a.w += v.y;
                    }
                    s_acc[4 * q] += a.x; s_acc[4 * q + 1] += a.y; s_acc[4 * q + 2] += a.z; s_acc[4 * q + 3] += a.w;
                } else {
                    float2 a = make_float2(0.f, 0.f);
                    for (int r = 0; r < rows; ++r) { const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 808); a.x += u.x; a.y += u.y; }
                    s_acc[808] += a.x; s_acc[809] += a.y;
                }
            }
        } else if (!(BV & 2)) {
            for (int o = t; o < MC; o += blockDim.x) {
                float a0 = 0.f, a1 = 0.f;
                int r = 0;
                for (; r + 1 < rows; r += 2) { a0 += tile[r * MC + o]; a1 += tile[(r + 1) * MC + o]; }
                if (r < rows) a0 += tile[r * MC + o];
                s_acc[o] += a0 + a1;
            }
        }
        for (int j = t; j < rows * 3; j += blockDim.x) {
            const int ss = j / 3, a = j - 3 * ss;
            float acc = 0.f;
            for (int cc = 0; cc < C; ++cc) acc += s_gp[(ss * C + cc) * 3 + a];
            gangles[n0 * 3 + j] = acc;
        }
        __syncthreads();
    }
    if (!(BV & 128)) for (int o = t; o < MC; o += blockDim.x) partial[int64_t(blockIdx.x) * MC + o] = s_acc[o];
}

template <int BV, int MINB>
static int launch_b(const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial, int64_t N, int S, int grid, cudaStream_t st) {
    const size_t smem = size_t(S * MC + S * TS + S * C * 3 + 4 + MC + 2 + MC) * 4;
    cudaFuncSetAttribute(bwd_kernel<BV, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    bwd_kernel<BV, MINB><<<grid, S * C, smem, st>>>(angles, spectrum, gout, gangles, partial, N, S, (N + S - 1) / S);
    return int(cudaGetLastError());
}

extern "C" int exp_wigner_bwd(int var, const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                              int64_t N, int S, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (var) {
        case 0: return launch_b<0, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 1: return launch_b<1, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 2: return launch_b<2, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 3: return launch_b<3, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 4: return launch_b<4, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 8: return launch_b<8, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 16: return launch_b<16, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 32: return launch_b<32, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 48: return launch_b<48, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 129: return launch_b<129, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 160: return launch_b<160, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 128: return launch_b<128, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 96: return launch_b<96, 3>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 100: return launch_b<0, 4>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        default: return -1;
    }
}

// ===================================================================== warp-private backward
// Each warp owns 3 samples x 10 channels (30 lanes) per iteration: private tile / trig / accumulators,
// no CTA barriers in the loop.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(uint32_t(__cvta_generic_to_shared(smem_dst))), "l"(gsrc) : "memory");
}
constexpr int SW = 3;                       // samples per warp
constexpr int WT_TILE = SW * MC;            // 2430 floats
constexpr int WT_FLOATS = WT_TILE + 2 + SW * TS + 4 + 96 + MC + 2;   // tile, trig, gp, acc  (~14.2 KB)

template <int BV>
__global__ void __launch_bounds__(160, 3)
bwd_warp_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
                float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int64_t nwt) {
    extern __shared__ __align__(16) float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    float* s_item = smem;                                   // [MC] shared by the CTA
    float* wbase = smem + MC + 2 + warp * WT_FLOATS;
    float* tile = wbase;
    float* s_trig = tile + WT_TILE + 2;
    float* s_gp = s_trig + SW * TS + 4;
    float* s_acc = s_gp + 96;
    for (int o = threadIdx.x; o < MC; o += blockDim.x) s_item[o] = __ldg(spectrum + o);
    for (int o = lane; o < MC; o += 32) s_acc[o] = 0.f;
    __syncthreads();
    const int s = lane / C, c = lane - s * C;
    const int64_t gw = int64_t(blockIdx.x) * nwarps + warp, tw = int64_t(gridDim.x) * nwarps;
    for (int64_t wt = gw; wt < nwt; wt += tw) {
        const int64_t n0 = wt * SW;
        const int rows = int(min(int64_t(SW), N - n0));
        const float* gsrc = gout + n0 * MC;
        const int n8 = rows * MC / 2;
        for (int i = lane; i < n8; i += 32) cp_async8(tile + 2 * i, gsrc + 2 * i);
        if (lane < rows * 3) {
            const int ss = lane / 3, a = lane - 3 * ss;
            float s1, c1;
            sincosf(__ldg(angles + n0 * 3 + lane), &s1, &c1);
            float2* dst = reinterpret_cast<float2*>(s_trig + ss * TS + a * 16);
            float cm = c1, sm = s1;
#pragma unroll
            for (int m = 1; m <= LMAX; ++m) { dst[m - 1] = make_float2(cm, sm); const float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn; }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (s < rows) {
            const float2* tg = reinterpret_cast<const float2*>(s_trig + s * TS);
            float ga = 0.f, gb = 0.f, gc = 0.f;
            degrees_bwd_from<0, BV>(s_item + c, tile + s * MC + c, tg, ga, gb, gc);
            s_gp[lane * 3 + 0] = ga; s_gp[lane * 3 + 1] = gb; s_gp[lane * 3 + 2] = gc;
        }
        __syncwarp();
        for (int q = lane; q < MC / 2; q += 32) {
            float2 a = *reinterpret_cast<float2*>(s_acc + 2 * q);
            for (int r = 0; r < rows; ++r) { const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 2 * q); a.x += u.x; a.y += u.y; }
            *reinterpret_cast<float2*>(s_acc + 2 * q) = a;
        }
        if (lane < rows * 3) {
            const int ss = lane / 3, a = lane - 3 * ss;
            float acc = 0.f;
            for (int cc = 0; cc < C; ++cc) acc += s_gp[(ss * C + cc) * 3 + a];
            gangles[n0 * 3 + lane] = acc;
        }
        __syncwarp();
    }
    // CTA-level deterministic sum of the warps' accumulators -> one partial row per CTA
    __syncthreads();
    for (int o = threadIdx.x; o < MC; o += blockDim.x) {
        float a = 0.f;
        for (int w = 0; w < nwarps; ++w) a += smem[MC + 2 + w * WT_FLOATS + WT_TILE + 2 + SW * TS + 4 + 96 + o];
        partial[int64_t(blockIdx.x) * MC + o] = a;
    }
}

extern "C" int exp_wigner_bwd_warp(int var, const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                                   int64_t N, int grid, int threads, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = size_t(MC + 2 + (threads / 32) * WT_FLOATS) * 4;
    cudaFuncSetAttribute(bwd_warp_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    bwd_warp_kernel<0><<<grid, threads, smem, st>>>(angles, spectrum, gout, gangles, partial, N, (N + SW - 1) / SW);
    return int(cudaGetLastError());
}

// ===================================================================== packed-f32x2 backward (two channels per thread)
using lv::wg::f32x2_t;
using lv::wg::vmul;
using lv::wg::vfma;
using lv::wg::vneg;
__device__ __forceinline__ f32x2_t vfma_vv(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2_t vmul_vv(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float2 vunpack(f32x2_t v) { float2 r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }

template <int L, bool TRANSPOSED>
__device__ __forceinline__ void xrot2(f32x2_t (&x)[2 * L + 1], const float2* __restrict__ cs) {
    const float4* cs4 = reinterpret_cast<const float4*>(cs);
#pragma unroll
    for (int p = 0; p < (L + 1) / 2; ++p) {
        float4 t = cs4[p];
        if (TRANSPOSED) { t.y = -t.y; t.w = -t.w; }
        {
            const int m = 2 * p + 1;
            const f32x2_t a = x[L - m], b = x[L + m];
            x[L - m] = vfma(t.x, a, vmul(t.y, b));
            x[L + m] = vfma(t.x, b, vmul(-t.y, a));
        }
        if (2 * p + 2 <= L) {
            const int m = 2 * p + 2;
            const f32x2_t a = x[L - m], b = x[L + m];
            x[L - m] = vfma(t.z, a, vmul(t.w, b));
            x[L + m] = vfma(t.z, b, vmul(-t.w, a));
        }
    }
}
// acc_p += m h[l-m] w[l+m], acc_n += m h[l+m] w[l-m]   (<h, G w> = acc_p - acc_n)
template <int L>
__device__ __forceinline__ void gdot2(const f32x2_t (&h)[2 * L + 1], const f32x2_t (&w)[2 * L + 1], f32x2_t& ap, f32x2_t& an) {
#pragma unroll
    for (int m = 1; m <= L; ++m) {
        ap = vfma_vv(m == 1 ? h[L - m] : vmul(float(m), h[L - m]), w[L + m], ap);
        an = vfma_vv(m == 1 ? h[L + m] : vmul(float(m), h[L + m]), w[L - m], an);
    }
}

constexpr int CP = C / 2;    // channel pairs per sample = stride of a column in 64-bit units
template <int L>
__device__ __forceinline__ void degree_bwd2(const f32x2_t* src, f32x2_t* g, const float2* tg, f32x2_t (&acc)[6]) {
    f32x2_t x[2 * L + 1], y[2 * L + 1], w2[2 * L + 1];
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = src[i * CP];
    xrot2<L, false>(x, tg + 16);
    jmul<L>(x, w2);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = w2[i];
    xrot2<L, false>(y, tg + 8);
    jmul<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = g[i * CP];
    xrot2<L, true>(y, tg);
    gdot2<L>(y, x, acc[0], acc[1]);
    jmul<L>(y, x);
    xrot2<L, true>(x, tg + 8);
    gdot2<L>(x, w2, acc[2], acc[3]);
    jmul<L>(x, y);
    xrot2<L, true>(y, tg + 16);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = src[i * CP];
    gdot2<L>(y, x, acc[4], acc[5]);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) g[i * CP] = y[i];
}
template <int L>
__device__ __forceinline__ void degrees_bwd2_from(const f32x2_t* srow, f32x2_t* trow, const float2* tg, f32x2_t (&acc)[6]) {
    degree_bwd2<L>(srow + L * L * CP, trow + L * L * CP, tg, acc);
    if constexpr (L < LMAX) degrees_bwd2_from<L + 1>(srow, trow, tg, acc);
}

// S samples per CTA, S*CP threads.  smem: tile [S][MC], trig [S][52], gp [S*CP][3], item [MC] (optional)
template <int ITEM_SMEM, int MINB>
__global__ void __launch_bounds__(160, MINB)
bwd2_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
            float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int S, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float* s_trig = tile + S * MC;
    float* s_gp = s_trig + S * TS;
    float* s_item = s_gp + S * CP * 3 + 4;
    const int t = threadIdx.x;
    const int s = t / CP, p = t - s * CP;
    if (ITEM_SMEM & 1) for (int o = t; o < MC; o += blockDim.x) s_item[o] = __ldg(spectrum + o);
    // per-thread column accumulators: quads q = t, t + blockDim, ... (MC = 810 -> 203 quads, <= 2 per thread at 160 threads)
    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
        const int64_t n0 = tile_idx * S;
        const int rows = int(min(int64_t(S), N - n0));
        if (!(ITEM_SMEM & 2)) tile_g2s(tile, gout + n0 * MC, rows * MC);
        stage_trig(s_trig, angles, n0, rows);
        tile_async_wait();
        __syncthreads();
        if (s < rows) {
            const float2* tg = reinterpret_cast<const float2*>(s_trig + s * TS);
            f32x2_t* trow = reinterpret_cast<f32x2_t*>(tile + s * MC) + p;
            const f32x2_t* srow = reinterpret_cast<const f32x2_t*>((ITEM_SMEM & 1) ? s_item : spectrum) + p;
            f32x2_t acc[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
            degrees_bwd2_from<0>(srow, trow, tg, acc);
            const float2 a0 = vunpack(acc[0]), a1 = vunpack(acc[1]), b0 = vunpack(acc[2]), b1 = vunpack(acc[3]), c0 = vunpack(acc[4]), c1 = vunpack(acc[5]);
            s_gp[t * 3 + 0] = (a0.x - a1.x) + (a0.y - a1.y);
            s_gp[t * 3 + 1] = (b0.x - b1.x) + (b0.y - b1.y);
            s_gp[t * 3 + 2] = (c0.x - c1.x) + (c0.y - c1.y);
        }
        __syncthreads();
        if (!(ITEM_SMEM & 4)) {
            // column sums into registers: thread owns quads t and t + blockDim.x
            for (int k = 0; k < 2; ++k) {
                const int q = t + k * blockDim.x;
                if (q < 203) {
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    const bool full = q < 202;
                    for (int r = 0; r < rows; ++r) {
                        const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q);
                        a.x += u.x; a.y += u.y;
                        if (full) { const float2 v = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q + 2); a.z += v.x; a.w += v.y; }
                    }
                    if (k == 0) { acc_a.x += a.x; acc_a.y += a.y; acc_a.z += a.z; acc_a.w += a.w; }
                    else { acc_b.x += a.x; acc_b.y += a.y; acc_b.z += a.z; acc_b.w += a.w; }
                }
            }
        }
        for (int j = t; j < rows * 3; j += blockDim.x) {
            const int ss = j / 3, a = j - 3 * ss;
            float acc = 0.f;
            for (int cc = 0; cc < CP; ++cc) acc += s_gp[(ss * CP + cc) * 3 + a];
            gangles[n0 * 3 + j] = acc;
        }
        __syncthreads();
    }
    float* prow = partial + int64_t(blockIdx.x) * MC;
    if (t < 203) { prow[4 * t] = acc_a.x; prow[4 * t + 1] = acc_a.y; if (t < 202) { prow[4 * t + 2] = acc_a.z; prow[4 * t + 3] = acc_a.w; } }
    const int q2 = t + blockDim.x;
    if (q2 < 203) { prow[4 * q2] = acc_b.x; prow[4 * q2 + 1] = acc_b.y; if (q2 < 202) { prow[4 * q2 + 2] = acc_b.z; prow[4 * q2 + 3] = acc_b.w; } }
}

template <int ITEM_SMEM, int MINB>
static int launch_b2(const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial, int64_t N, int S, int grid, cudaStream_t st) {
    const size_t smem = size_t(S * MC + S * TS + S * CP * 3 + 4 + ((ITEM_SMEM & 1) ? MC : 0)) * 4;
    cudaError_t e = cudaFuncSetAttribute(bwd2_kernel<ITEM_SMEM, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    bwd2_kernel<ITEM_SMEM, MINB><<<grid, S * CP, smem, st>>>(angles, spectrum, gout, gangles, partial, N, S, (N + S - 1) / S);
    return int(cudaGetLastError());
}
extern "C" int exp_wigner_bwd2(int var, const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                               int64_t N, int S, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    switch (var) {
        case 0: return launch_b2<0, 2>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 1: return launch_b2<1, 1>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 2: return launch_b2<0, 1>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        case 6: return launch_b2<6, 2>(angles, spectrum, gout, gangles, partial, N, S, grid, st);
        default: return -1;
    }
}

// ===================================================================== channel-major backward
// warp = one channel, lane = one of 32 samples: trig loads are dense (no 10x broadcast waste), item loads are
// warp-uniform, tile accesses are 2-way conflicted (row stride 810 words) but serve 32 samples each.
constexpr int CM_S = 32;
template <int L, bool TRANSPOSED>
__device__ __forceinline__ void xrot_cm(float (&x)[2 * L + 1], const float4* __restrict__ cs4) {
#pragma unroll
    for (int p = 0; p < (L + 1) / 2; ++p) {
        float4 t = cs4[p * CM_S];
        if (TRANSPOSED) { t.y = -t.y; t.w = -t.w; }
        { const int m = 2 * p + 1; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.x, a, t.y * b); x[L + m] = fmaf(t.x, b, -(t.y * a)); }
        if (2 * p + 2 <= L) { const int m = 2 * p + 2; const float a = x[L - m], b = x[L + m]; x[L - m] = fmaf(t.z, a, t.w * b); x[L + m] = fmaf(t.z, b, -(t.w * a)); }
    }
}
template <int L>
__device__ __forceinline__ void degree_bwd_cm(const float* __restrict__ src, float* g, const float4* tg, float& ga, float& gb, float& gc) {
    float x[2 * L + 1], y[2 * L + 1], w2[2 * L + 1];
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = __ldg(src + i * C);
    xrot_cm<L, false>(x, tg + 8 * CM_S);
    jmul<L>(x, w2);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = w2[i];
    xrot_cm<L, false>(y, tg + 4 * CM_S);
    jmul<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = g[i * C];
    xrot_cm<L, true>(y, tg);
    ga += gdot<L>(y, x);
    jmul<L>(y, x);
    xrot_cm<L, true>(x, tg + 4 * CM_S);
    gb += gdot<L>(x, w2);
    jmul<L>(x, y);
    xrot_cm<L, true>(y, tg + 8 * CM_S);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = __ldg(src + i * C);
    gc += gdot<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) g[i * C] = y[i];
}
template <int L>
__device__ __forceinline__ void degrees_bwd_cm_from(const float* srow, float* trow, const float4* tg, float& ga, float& gb, float& gc) {
    degree_bwd_cm<L>(srow + L * L * C, trow + L * L * C, tg, ga, gb, gc);
    if constexpr (L < LMAX) degrees_bwd_cm_from<L + 1>(srow, trow, tg, ga, gb, gc);
}

__global__ void __launch_bounds__(320, 2)
bwdcm_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
             float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;                                             // [32][810]
    float4* s_trig = reinterpret_cast<float4*>(tile + CM_S * MC);    // [3][4][32] float4
    float* s_gp = reinterpret_cast<float*>(s_trig + 12 * CM_S);      // [10][32][3]
    const int t = threadIdx.x, c = t >> 5, s = t & 31;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
        const int64_t n0 = tile_idx * CM_S;
        const int rows = int(min(int64_t(CM_S), N - n0));
        tile_g2s(tile, gout + n0 * MC, rows * MC);
        if (t < rows * 3) {
            const int ss = t / 3, a = t - 3 * ss;
            float s1, c1;
            sincosf(__ldg(angles + n0 * 3 + t), &s1, &c1);
            float cm = c1, sm = s1;
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float4 v;
                v.x = cm; v.y = sm;
                float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn;
                v.z = cm; v.w = sm;
                cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn;
                s_trig[(a * 4 + p) * CM_S + ss] = v;
            }
        }
        tile_async_wait();
        __syncthreads();
        if (s < rows) {
            float ga = 0.f, gb = 0.f, gc = 0.f;
            degrees_bwd_cm_from<0>(spectrum + c, tile + s * MC + c, s_trig + s, ga, gb, gc);
            s_gp[(c * CM_S + s) * 3 + 0] = ga; s_gp[(c * CM_S + s) * 3 + 1] = gb; s_gp[(c * CM_S + s) * 3 + 2] = gc;
        }
        __syncthreads();
        if (t < 203) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            const bool full = t < 202;
            for (int r = 0; r < rows; ++r) {
                const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 4 * t);
                a.x += u.x; a.y += u.y;
                if (full) { const float2 v = *reinterpret_cast<const float2*>(tile + r * MC + 4 * t + 2); a.z += v.x; a.w += v.y; }
            }
            acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
        } else if (t >= 224 && t < 224 + rows * 3) {
            const int j = t - 224, ss = j / 3, a = j - 3 * ss;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) sum += s_gp[(cc * CM_S + ss) * 3 + a];
            gangles[n0 * 3 + j] = sum;
        }
        __syncthreads();
    }
    float* prow = partial + int64_t(blockIdx.x) * MC;
    if (t < 203) { prow[4 * t] = acc.x; prow[4 * t + 1] = acc.y; if (t < 202) { prow[4 * t + 2] = acc.z; prow[4 * t + 3] = acc.w; } }
}

extern "C" int exp_wigner_bwdcm(const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                                int64_t N, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = size_t(CM_S * MC) * 4 + 12 * CM_S * 16 + C * CM_S * 3 * 4;
    cudaError_t e = cudaFuncSetAttribute(bwdcm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    bwdcm_kernel<<<grid, 320, smem, st>>>(angles, spectrum, gout, gangles, partial, N, (N + CM_S - 1) / CM_S);
    return int(cudaGetLastError());
}

// ===================================================================== packed + channel-major backward
// warp = one channel PAIR (f32x2), lane = sample; CTA = 5 warps x 32 samples.
template <int L, bool TRANSPOSED>
__device__ __forceinline__ void xrot2_cm(f32x2_t (&x)[2 * L + 1], const float4* __restrict__ cs4) {
#pragma unroll
    for (int p = 0; p < (L + 1) / 2; ++p) {
        float4 t = cs4[p * CM_S];
        if (TRANSPOSED) { t.y = -t.y; t.w = -t.w; }
        { const int m = 2 * p + 1; const f32x2_t a = x[L - m], b = x[L + m]; x[L - m] = vfma(t.x, a, vmul(t.y, b)); x[L + m] = vfma(t.x, b, vmul(-t.y, a)); }
        if (2 * p + 2 <= L) { const int m = 2 * p + 2; const f32x2_t a = x[L - m], b = x[L + m]; x[L - m] = vfma(t.z, a, vmul(t.w, b)); x[L + m] = vfma(t.z, b, vmul(-t.w, a)); }
    }
}
template <int L>
__device__ __forceinline__ void degree_bwd2_cm(const f32x2_t* __restrict__ src, f32x2_t* g, const float4* tg, f32x2_t (&acc)[6]) {
    f32x2_t x[2 * L + 1], y[2 * L + 1], w2[2 * L + 1];
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = __ldg(src + i * CP);
    xrot2_cm<L, false>(x, tg + 8 * CM_S);
    jmul<L>(x, w2);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = w2[i];
    xrot2_cm<L, false>(y, tg + 4 * CM_S);
    jmul<L>(y, x);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) y[i] = g[i * CP];
    xrot2_cm<L, true>(y, tg);
    gdot2<L>(y, x, acc[0], acc[1]);
    jmul<L>(y, x);
    xrot2_cm<L, true>(x, tg + 4 * CM_S);
    gdot2<L>(x, w2, acc[2], acc[3]);
    jmul<L>(x, y);
    xrot2_cm<L, true>(y, tg + 8 * CM_S);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) x[i] = __ldg(src + i * CP);
    gdot2<L>(y, x, acc[4], acc[5]);
#pragma unroll
    for (int i = 0; i < 2 * L + 1; ++i) g[i * CP] = y[i];
}
template <int L>
__device__ __forceinline__ void degrees_bwd2_cm_from(const f32x2_t* srow, f32x2_t* trow, const float4* tg, f32x2_t (&acc)[6]) {
    degree_bwd2_cm<L>(srow + L * L * CP, trow + L * L * CP, tg, acc);
    if constexpr (L < LMAX) degrees_bwd2_cm_from<L + 1>(srow, trow, tg, acc);
}

__global__ void __launch_bounds__(160, 2)
bwd2cm_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
              float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* tile = smem;
    float4* s_trig = reinterpret_cast<float4*>(tile + CM_S * MC);
    float* s_gp = reinterpret_cast<float*>(s_trig + 12 * CM_S);      // [5][32][3]
    const int t = threadIdx.x, p = t >> 5, s = t & 31;
    float4 acc_a = make_float4(0.f, 0.f, 0.f, 0.f), acc_b = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t tile_idx = blockIdx.x; tile_idx < ntiles; tile_idx += gridDim.x) {
        const int64_t n0 = tile_idx * CM_S;
        const int rows = int(min(int64_t(CM_S), N - n0));
        tile_g2s(tile, gout + n0 * MC, rows * MC);
        if (t < rows * 3) {
            const int ss = t / 3, a = t - 3 * ss;
            float s1, c1;
            sincosf(__ldg(angles + n0 * 3 + t), &s1, &c1);
            float cm = c1, sm = s1;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float4 v;
                v.x = cm; v.y = sm;
                float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn;
                v.z = cm; v.w = sm;
                cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn;
                s_trig[(a * 4 + q) * CM_S + ss] = v;
            }
        }
        tile_async_wait();
        __syncthreads();
        if (s < rows) {
            f32x2_t acc[6] = {0ull, 0ull, 0ull, 0ull, 0ull, 0ull};
            degrees_bwd2_cm_from<0>(reinterpret_cast<const f32x2_t*>(spectrum) + p, reinterpret_cast<f32x2_t*>(tile + s * MC) + p, s_trig + s, acc);
            const float2 a0 = vunpack(acc[0]), a1 = vunpack(acc[1]), b0 = vunpack(acc[2]), b1 = vunpack(acc[3]), c0 = vunpack(acc[4]), c1 = vunpack(acc[5]);
            s_gp[(p * CM_S + s) * 3 + 0] = (a0.x - a1.x) + (a0.y - a1.y);
            s_gp[(p * CM_S + s) * 3 + 1] = (b0.x - b1.x) + (b0.y - b1.y);
            s_gp[(p * CM_S + s) * 3 + 2] = (c0.x - c1.x) + (c0.y - c1.y);
        }
        __syncthreads();
        for (int k = 0; k < 2; ++k) {
            const int q = t + k * 160;
            if (q < 203) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                const bool full = q < 202;
                for (int r = 0; r < rows; ++r) {
                    const float2 u = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q);
                    a.x += u.x; a.y += u.y;
                    if (full) { const float2 v = *reinterpret_cast<const float2*>(tile + r * MC + 4 * q + 2); a.z += v.x; a.w += v.y; }
                }
                if (k == 0) { acc_a.x += a.x; acc_a.y += a.y; acc_a.z += a.z; acc_a.w += a.w; }
                else { acc_b.x += a.x; acc_b.y += a.y; acc_b.z += a.z; acc_b.w += a.w; }
            }
        }
        if (t >= 64 && t < 64 + rows * 3) {
            const int j = t - 64, ss = j / 3, a = j - 3 * ss;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < CP; ++cc) sum += s_gp[(cc * CM_S + ss) * 3 + a];
            gangles[n0 * 3 + j] = sum;
        }
        __syncthreads();
    }
    float* prow = partial + int64_t(blockIdx.x) * MC;
    if (t < 203) { prow[4 * t] = acc_a.x; prow[4 * t + 1] = acc_a.y; if (t < 202) { prow[4 * t + 2] = acc_a.z; prow[4 * t + 3] = acc_a.w; } }
    const int q2 = t + 160;
    if (q2 < 203) { prow[4 * q2] = acc_b.x; prow[4 * q2 + 1] = acc_b.y; if (q2 < 202) { prow[4 * q2 + 2] = acc_b.z; prow[4 * q2 + 3] = acc_b.w; } }
}

extern "C" int exp_wigner_bwd2cm(const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                                 int64_t N, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = size_t(CM_S * MC) * 4 + 12 * CM_S * 16 + CP * CM_S * 3 * 4;
    cudaError_t e = cudaFuncSetAttribute(bwd2cm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    bwd2cm_kernel<<<grid, 160, smem, st>>>(angles, spectrum, gout, gangles, partial, N, (N + CM_S - 1) / CM_S);
    return int(cudaGetLastError());
}

// ===================================================================== warp-specialised, degree-pipelined backward
// warps 0-4: math (thread = (sample, channel), 16 samples x 10 channels); warp 5: service warp that, per degree,
// sums the finished gradient rows over the 16 samples and immediately refills them with the next tile's rows.
// The math warps never wait for a whole-tile load or a reduce phase.
constexpr int WS_S = 16;
constexpr int WS_MATH = 160;
constexpr int WS_NSVC = 2;                 // service warps
constexpr int WS_THREADS = WS_MATH + 32 * WS_NSVC;
constexpr int WS_SL = 32 * WS_NSVC;        // service lanes

__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_cp_async_arrive(uint64_t* bar) { asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct WsShared {
    float* tile; float* trig; float* gp; float* acc; float* item; uint64_t* g_ready; uint64_t* trig_ready;
};

template <int L>
__device__ __forceinline__ void ws_math_degrees(const WsShared& sh, const float* srow, float* trow, const float2* tg, uint32_t parity,
                                                float& ga, float& gb, float& gc) {
    mbar_wait(sh.g_ready + L, parity);                       // rows of degree L of this tile have landed
    degree_bwd<L, 0>(srow + L * L * C, trow + L * L * C, tg, ga, gb, gc);
    bar_arrive(1 + L, WS_THREADS);                            // gradient rows of degree L are final
    if constexpr (L < LMAX) ws_math_degrees<L + 1>(sh, srow, trow, tg, parity, ga, gb, gc);
}

// service warp: issue the loads of the degree-L rows of tile at sample n0 (rows samples), 8-byte pieces
template <int L>
__device__ __forceinline__ void ws_load_degree(const WsShared& sh, const float* __restrict__ gout, int64_t n0, int rows, int lane) {
    constexpr int PPS = (2 * L + 1) * (C / 2);                // 8-byte pieces per sample for this degree
    const int total = rows * PPS;
    for (int j = lane; j < total; j += WS_SL) {
        const int s = j / PPS, k = j - s * PPS;
        const int off = s * MC + L * L * C + 2 * k;
        cp_async8(sh.tile + off, gout + n0 * MC + off);
    }
    mbar_cp_async_arrive(sh.g_ready + L);
}
template <int L>
__device__ __forceinline__ void ws_service_degrees(const WsShared& sh, const float* __restrict__ gout, int rows, bool has_next, int64_t n_next,
                                                   int rows_next, int lane) {
    bar_sync(1 + L, WS_THREADS);                              // all math threads finished degree L
    constexpr int NP = (2 * L + 1) * (C / 2);                 // column pairs of this degree
    for (int p = lane; p < NP; p += WS_SL) {
        const int col = L * L * C + 2 * p;
        float2 a = *reinterpret_cast<float2*>(sh.acc + col);
        for (int r = 0; r < rows; ++r) { const float2 u = *reinterpret_cast<const float2*>(sh.tile + r * MC + col); a.x += u.x; a.y += u.y; }
        *reinterpret_cast<float2*>(sh.acc + col) = a;
    }
    bar_sync(12, WS_SL);                                      // all service lanes have summed this degree
    if (has_next) ws_load_degree<L>(sh, gout, n_next, rows_next, lane);
    if constexpr (L < LMAX) ws_service_degrees<L + 1>(sh, gout, rows, has_next, n_next, rows_next, lane);
}
template <int L>
__device__ __forceinline__ void ws_load_all(const WsShared& sh, const float* __restrict__ gout, int64_t n0, int rows, int lane) {
    ws_load_degree<L>(sh, gout, n0, rows, lane);
    if constexpr (L < LMAX) ws_load_all<L + 1>(sh, gout, n0, rows, lane);
}
__device__ __forceinline__ void ws_trig(float* s_trig, const float* __restrict__ angles, int64_t n0, int rows, int lane) {
    for (int j = lane; j < rows * 3; j += WS_SL) {
        const int s = j / 3, a = j - 3 * s;
        float s1, c1;
        sincosf(__ldg(angles + n0 * 3 + j), &s1, &c1);
        float2* dst = reinterpret_cast<float2*>(s_trig + s * TS + a * 16);
        float cm = c1, sm = s1;
#pragma unroll
        for (int m = 1; m <= LMAX; ++m) { dst[m - 1] = make_float2(cm, sm); const float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn; }
    }
}

__global__ void __launch_bounds__(WS_THREADS, 3)
bwdws_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
             float* __restrict__ gangles, float* __restrict__ partial, int64_t N, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    WsShared sh;
    sh.tile = smem;                                  // [16][810]
    sh.trig = sh.tile + WS_S * MC;                   // [2][16][52]
    sh.gp = sh.trig + 2 * WS_S * TS;                 // [160][3]
    sh.acc = sh.gp + 2 * WS_MATH * 3;                // [810] (+2 pad)
    sh.item = sh.acc + MC + 2;                       // [810] (+2 pad)
    sh.g_ready = reinterpret_cast<uint64_t*>(sh.item + MC + 2);   // [9]
    sh.trig_ready = sh.g_ready + 9;                  // [1]
    const int t = threadIdx.x, warp = t >> 5, lane = t - WS_MATH;    // lane: index among the service lanes
    for (int o = t; o < MC; o += blockDim.x) { sh.acc[o] = 0.f; sh.item[o] = __ldg(spectrum + o); }
    if (t == 0) {
        for (int l = 0; l <= LMAX; ++l) mbar_init(sh.g_ready + l, WS_SL);
        mbar_init(sh.trig_ready, WS_SL);
    }
    __syncthreads();
    const int64_t first = blockIdx.x, stride = gridDim.x;
    if (warp < 5) {
        // ---------------------------------------------------------------- math warps
        const int s = t / C, c = t - s * C;
        uint32_t it = 0;
        for (int64_t tile_idx = first; tile_idx < ntiles; tile_idx += stride, ++it) {
            const int64_t n0 = tile_idx * WS_S;
            const int rows = int(min(int64_t(WS_S), N - n0));
            const uint32_t parity = it & 1u;
            mbar_wait(sh.trig_ready, parity);
            float ga = 0.f, gb = 0.f, gc = 0.f;
            if (s < rows) {
                const float2* tg = reinterpret_cast<const float2*>(sh.trig + (it & 1u) * WS_S * TS + s * TS);
                ws_math_degrees<0>(sh, sh.item + c, sh.tile + s * MC + c, tg, parity, ga, gb, gc);
            } else {
                // idle column (partial last tile): still take part in the per-degree hand-shake
#pragma unroll
                for (int l = 0; l <= LMAX; ++l) { mbar_wait(sh.g_ready + l, parity); bar_arrive(1 + l, WS_THREADS); }
            }
            float* gp = sh.gp + (it & 1u) * WS_MATH * 3;     // double-buffered: the service warp reads it a little later
            gp[t * 3 + 0] = ga; gp[t * 3 + 1] = gb; gp[t * 3 + 2] = gc;
            bar_arrive(10 + (it & 1u), WS_THREADS);          // angle-gradient parts are written
        }
    } else {
        // ---------------------------------------------------------------- service warp
        uint32_t it = 0;
        if (first < ntiles) {
            const int rows0 = int(min(int64_t(WS_S), N - first * WS_S));
            ws_trig(sh.trig, angles, first * WS_S, rows0, lane);
            mbar_arrive(sh.trig_ready);
            ws_load_all<0>(sh, gout, first * WS_S, rows0, lane);
        }
        for (int64_t tile_idx = first; tile_idx < ntiles; tile_idx += stride, ++it) {
            const int64_t n0 = tile_idx * WS_S;
            const int rows = int(min(int64_t(WS_S), N - n0));
            const int64_t nxt = tile_idx + stride;
            const bool has_next = nxt < ntiles;
            const int64_t n_next = nxt * WS_S;
            const int rows_next = has_next ? int(min(int64_t(WS_S), N - n_next)) : 0;
            if (has_next) {
                ws_trig(sh.trig + ((it + 1) & 1u) * WS_S * TS, angles, n_next, rows_next, lane);
                mbar_arrive(sh.trig_ready);
            }
            ws_service_degrees<0>(sh, gout, rows, has_next, n_next, rows_next, lane);
            bar_sync(10 + (it & 1u), WS_THREADS);
            const float* gp = sh.gp + (it & 1u) * WS_MATH * 3;
            for (int j = lane; j < rows * 3; j += WS_SL) {
                const int ss = j / 3, a = j - 3 * ss;
                float sum = 0.f;
#pragma unroll
                for (int cc = 0; cc < C; ++cc) sum += gp[(ss * C + cc) * 3 + a];
                gangles[n0 * 3 + j] = sum;
            }
        }
    }
    __syncthreads();
    for (int o = t; o < MC; o += blockDim.x) partial[int64_t(blockIdx.x) * MC + o] = sh.acc[o];
}

extern "C" int exp_wigner_bwdws(const float* angles, const float* spectrum, const float* gout, float* gangles, float* partial,
                                int64_t N, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = size_t(WS_S * MC + 2 * WS_S * TS + 2 * WS_MATH * 3 + MC + 2 + MC + 2) * 4 + 10 * 8;
    cudaError_t e = cudaFuncSetAttribute(bwdws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    bwdws_kernel<<<grid, WS_THREADS, smem, st>>>(angles, spectrum, gout, gangles, partial, N, (N + WS_S - 1) / WS_S);
    return int(cudaGetLastError());
}

// ===================================================================== one persistent CTA per SM, 3 math groups, 4 tile buffers
// Tile j of this CTA lives in buffer j % 4 and is processed by group j % 3.  A group that finishes a tile hands it to the
// copy engine twice: a TMA bulk reduce-add of the gradient rows into the group's fp32 accumulator in global memory (L2),
// then a TMA bulk load of tile j + 4 into the same buffer.  Three tiles are always being computed while the fourth loads.
constexpr int Q_GROUPS = 3, Q_BUFS = 4, Q_GT = 160, Q_THREADS = Q_GROUPS * Q_GT;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(Q_THREADS, 1)
bwdq_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
            float* __restrict__ gangles, float* __restrict__ gacc, int64_t N, int64_t ntiles) {
    extern __shared__ __align__(16) float smem[];
    float* tiles = smem;                                          // [4][16][810]
    float* trig_all = tiles + Q_BUFS * WS_S * MC;                  // [3][16][52]
    float* gp_all = trig_all + Q_GROUPS * WS_S * TS;               // [3][160][3]
    uint64_t* full = reinterpret_cast<uint64_t*>(gp_all + Q_GROUPS * Q_GT * 3);   // [4]
    const int tid = threadIdx.x, g = tid / Q_GT, t = tid - g * Q_GT;
    const int s = t / C, c = t - s * C;
    float* s_trig = trig_all + g * WS_S * TS;
    float* s_gp = gp_all + g * Q_GT * 3;
    // tiles of this CTA: blockIdx.x + j * gridDim.x, j = 0, 1, ...
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int64_t my_tiles = first < ntiles ? (ntiles - first + stride - 1) / stride : 0;
    if (tid == 0) {
        for (int b = 0; b < Q_BUFS; ++b) mbar_init(full + b, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int j = 0; j < Q_BUFS && j < my_tiles; ++j) {
            const int64_t n0 = (first + j * stride) * WS_S;
            const uint32_t bytes = uint32_t(min(int64_t(WS_S), N - n0)) * MC * 4u;
            mbar_expect_tx(full + j, bytes);
            bulk_load(tiles + j * WS_S * MC, gout + n0 * MC, bytes, full + j);
        }
    }
    for (int64_t j = g; j < my_tiles; j += Q_GROUPS) {
        const int buf = int(j % Q_BUFS);
        const uint32_t parity = uint32_t(j / Q_BUFS) & 1u;
        const int64_t n0 = (first + j * stride) * WS_S;
        const int rows = int(min(int64_t(WS_S), N - n0));
        float* tile = tiles + buf * WS_S * MC;
        // trig table of this tile (group-private), then wait for the tile itself
        for (int q = t; q < rows * 3; q += Q_GT) {
            const int ss = q / 3, a = q - 3 * ss;
            float s1, c1;
            sincosf(__ldg(angles + n0 * 3 + q), &s1, &c1);
            float2* dst = reinterpret_cast<float2*>(s_trig + ss * TS + a * 16);
            float cm = c1, sm = s1;
#pragma unroll
            for (int m = 1; m <= LMAX; ++m) { dst[m - 1] = make_float2(cm, sm); const float cn = fmaf(cm, c1, -(sm * s1)); sm = fmaf(sm, c1, cm * s1); cm = cn; }
        }
        mbar_wait(full + buf, parity);
        bar_sync(1 + g, Q_GT);
        if (s < rows) {
            float ga = 0.f, gb = 0.f, gc = 0.f;
            degrees_bwd_from<0, 8>(spectrum + c, tile + s * MC + c, reinterpret_cast<const float2*>(s_trig + s * TS), ga, gb, gc);
            s_gp[t * 3 + 0] = ga; s_gp[t * 3 + 1] = gb; s_gp[t * 3 + 2] = gc;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bar_sync(1 + g, Q_GT);
        if (t == 0) {
            const uint32_t bytes = uint32_t(rows) * MC * 4u;
            float* acc = gacc + (int64_t(blockIdx.x) * Q_GROUPS + g) * WS_S * MC;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                         :: "l"(acc), "r"(smem_u32(tile)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            const int64_t jn = j + Q_BUFS;
            if (jn < my_tiles) {
                const int64_t nn = (first + jn * stride) * WS_S;
                const uint32_t nb = uint32_t(min(int64_t(WS_S), N - nn)) * MC * 4u;
                mbar_expect_tx(full + buf, nb);
                bulk_load(tile, gout + nn * MC, nb, full + buf);
            }
        }
        for (int q = t; q < rows * 3; q += Q_GT) {
            const int ss = q / 3, a = q - 3 * ss;
            float sum = 0.f;
#pragma unroll
            for (int cc = 0; cc < C; ++cc) sum += s_gp[(ss * C + cc) * 3 + a];
            gangles[n0 * 3 + q] = sum;
        }
    }
    // all bulk reductions of this CTA must have completed before the kernel ends
    if (t == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int exp_wigner_bwdq(const float* angles, const float* spectrum, const float* gout, float* gangles, float* gacc,
                               int64_t N, int grid, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const size_t smem = size_t(Q_BUFS * WS_S * MC + Q_GROUPS * WS_S * TS + Q_GROUPS * Q_GT * 3) * 4 + Q_BUFS * 8;
    cudaError_t e = cudaFuncSetAttribute(bwdq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    e = cudaMemsetAsync(gacc, 0, size_t(grid) * Q_GROUPS * WS_S * MC * 4, st);
    if (e != cudaSuccess) return int(e);
    bwdq_kernel<<<grid, Q_THREADS, smem, st>>>(angles, spectrum, gout, gangles, gacc, N, (N + WS_S - 1) / WS_S);
    return int(cudaGetLastError());
}
