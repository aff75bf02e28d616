// Encoder heads fused into the SO(3) reparameterize kernels (sm_100a, FP32) -- SURVEY.md 8f-2.
//
// One kernel per direction replaces, for a batch of encoder features h (B, Din):
//   mean head     Linear(Din -> 3 / 4 / 5 / 6)  + rodrigues / quaternions_to_group_matrix / s2s1rodrigues / s2s2_gram_schmidt
//                 AlgebraMean reparameterize.py:148-155, QuaternionMean :158-164, S2S1Mean :167-181 (axis and (cos, sin)
//                 normalised first), S2S2Mean :184-197 (Gram-Schmidt in float64 and cast back, exactly as :195-197)
//   sigma head    softplus(Linear(Din -> 3))                        N0reparameterize reparameterize.py:117-121
//   sampling      v = eps * sigma, z = mu @ rodrigues(v), wrapped log-density, optionally matrix -> ZYZ Euler
//                 (the body of reparam.cu, shared through reparam_core.cuh)
// so that mu and sigma never make a round trip through HBM between five separate launches (two GEMMs, softplus, the
// mean map, the sampler) and, in the backward, the per-sample g_mu / g_sigma never exist in memory: the thread that
// owns a sample pulls them back through the mean map and softplus and produces its row of g_h; the weight and bias
// gradients (a (Dm+3) x (Din+1) matrix) are reduced over the CTA's samples in shared memory and over the CTAs by a
// second tiny kernel (deterministic, no atomics).
//
// The two heads arrive as their own Linear parameters: Wm (Dm, Din), bm (Dm), Ws (3, Din), bs (3); Din <= 32.
// Inside the kernel they are one ((Dm+3), Din+1) matrix [mean head; sigma head | bias], and so is the gradient gWb.
#include "common.cuh"
#include "reparam_core.cuh"

namespace lv {

constexpr int HR_TILE = 256, HR_WPAD = 12, HR_MAX_DIN = 32;
enum { HR_ALG = 0, HR_QUAT = 1, HR_S2S2 = 2, HR_S2S1 = 3 };
__host__ __device__ constexpr int hr_mean_rows(int mode) { return mode == HR_ALG ? 3 : mode == HR_QUAT ? 4 : mode == HR_S2S2 ? 6 : 5; }

// softplus with torch's defaults (beta = 1, threshold = 20) and its derivative
__device__ __forceinline__ float hr_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float hr_softplus_grad(float x) { return x > 20.f ? 1.f : 1.f / (1.f + expf(-x)); }

template <int MODE>
__device__ __forceinline__ void hr_mean_fwd(const float* pre, float (&m)[9]) {
    if (MODE == HR_ALG) rodrigues_fwd(pre, m);
    else if (MODE == HR_QUAT) quat_to_mat_fwd(pre, m);
    else if (MODE == HR_S2S1) {
        // S2S1Mean (reparameterize.py:167-181): unit axis pre[0:3] / |.|, unit (cos, sin) pre[3:5] / |.|, then s2s1rodrigues
        const float i2 = 1.f / sqrtf(pre[0] * pre[0] + pre[1] * pre[1] + pre[2] * pre[2]);
        const float i1 = 1.f / sqrtf(pre[3] * pre[3] + pre[4] * pre[4]);
        const float u[3] = {pre[0] * i2, pre[1] * i2, pre[2] * i2};
        axis_angle_matrix(u, pre[4] * i1, 1.f - pre[3] * i1, m);
    } else {
        double v1[3] = {pre[0], pre[1], pre[2]}, v2[3] = {pre[3], pre[4], pre[5]}, R[9];
        s2s2_fwd(v1, v2, R);
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = float(R[j]);
    }
}
template <int MODE>
__device__ __forceinline__ void hr_mean_bwd(const float* pre, const float (&gm)[9], float* gpre) {
    if (MODE == HR_ALG) rodrigues_bwd(pre, gm, gpre);
    else if (MODE == HR_QUAT) quat_to_mat_bwd(pre, gm, gpre);
    else if (MODE == HR_S2S1) {
        const float i2 = 1.f / sqrtf(pre[0] * pre[0] + pre[1] * pre[1] + pre[2] * pre[2]);
        const float i1 = 1.f / sqrtf(pre[3] * pre[3] + pre[4] * pre[4]);
        const float u[3] = {pre[0] * i2, pre[1] * i2, pre[2] * i2}, cs[2] = {pre[3] * i1, pre[4] * i1};
        float gu[3], gs, gw;
        axis_angle_matrix_bwd(u, cs[1], 1.f - cs[0], gm, gu, &gs, &gw);
        const float gcs[2] = {-gw, gs};
        // x -> x / |x|:  g_x = (g_e - e (e . g_e)) / |x|
        const float du = u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2], dc = cs[0] * gcs[0] + cs[1] * gcs[1];
#pragma unroll
        for (int j = 0; j < 3; ++j) gpre[j] = (gu[j] - u[j] * du) * i2;
#pragma unroll
        for (int j = 0; j < 2; ++j) gpre[3 + j] = (gcs[j] - cs[j] * dc) * i1;
    } else {
        double v1[3] = {pre[0], pre[1], pre[2]}, v2[3] = {pre[3], pre[4], pre[5]}, G[9], g1[3], g2[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) G[j] = gm[j];
        s2s2_bwd(v1, v2, G, g1, g2);
#pragma unroll
        for (int j = 0; j < 3; ++j) { gpre[j] = float(g1[j]); gpre[3 + j] = float(g2[j]); }
    }
}

// rows [i0, i0+rows) of the (n,B,Din) broadcast view of h (B,Din) -> smem, contiguous unless the tile wraps
__device__ __forceinline__ void hr_stage_h(float* __restrict__ dst, const float* __restrict__ h, int64_t i0, int rows, int64_t B, int Din) {
    const int64_t b0 = i0 < B ? i0 : i0 % B;
    if (b0 + rows <= B) {
        tile_g2s(dst, h + b0 * Din, rows * Din);
    } else {
        for (int idx = threadIdx.x; idx < rows * Din; idx += blockDim.x) {
            const int r = idx / Din, c = idx - r * Din;
            dst[idx] = __ldg(h + ((i0 + r) % B) * Din + c);
        }
    }
}
// [Wm; Ws] (row-major) and [bm; bs] -> s_w[(Din+1)][HR_WPAD]: column d of the stacked weight in row d, the bias in row Din
__device__ __forceinline__ void hr_stage_w(float* __restrict__ s_w, const float* __restrict__ Wm, const float* __restrict__ bm,
                                           const float* __restrict__ Ws, const float* __restrict__ bs, int DM, int Din) {
    for (int idx = threadIdx.x; idx < (Din + 1) * HR_WPAD; idx += blockDim.x) {
        const int d = idx / HR_WPAD, j = idx - d * HR_WPAD;
        float v = 0.f;
        if (j < DM) v = d < Din ? __ldg(Wm + j * Din + d) : __ldg(bm + j);
        else if (j < DM + 3) v = d < Din ? __ldg(Ws + (j - DM) * Din + d) : __ldg(bs + (j - DM));
        s_w[idx] = v;
    }
}
// pre = bias + W h_t   (DT <= 9 outputs; three broadcast 128-bit reads of W^T per input feature)
template <int DT>
__device__ __forceinline__ void hr_linear(const float* __restrict__ s_w, const float* __restrict__ hrow, int Din, float (&pre)[9]) {
    const float4* w4 = reinterpret_cast<const float4*>(s_w);
    float acc[12];
    {
        const float4 a = w4[Din * 3], b = w4[Din * 3 + 1], c = w4[Din * 3 + 2];
        acc[0] = a.x; acc[1] = a.y; acc[2] = a.z; acc[3] = a.w; acc[4] = b.x; acc[5] = b.y; acc[6] = b.z; acc[7] = b.w;
        acc[8] = c.x; acc[9] = c.y; acc[10] = c.z; acc[11] = c.w;
    }
    for (int d = 0; d < Din; ++d) {
        const float x = hrow[d];
        const float4 a = w4[d * 3], b = w4[d * 3 + 1], c = w4[d * 3 + 2];
        acc[0] = fmaf(a.x, x, acc[0]); acc[1] = fmaf(a.y, x, acc[1]); acc[2] = fmaf(a.z, x, acc[2]); acc[3] = fmaf(a.w, x, acc[3]);
        acc[4] = fmaf(b.x, x, acc[4]); acc[5] = fmaf(b.y, x, acc[5]); acc[6] = fmaf(b.z, x, acc[6]); acc[7] = fmaf(b.w, x, acc[7]);
        if (DT > 8) acc[8] = fmaf(c.x, x, acc[8]);
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) pre[j] = acc[j];
}

// ------------------------------------------------------------------ forward
template <int MODE, int KT, bool EULER>
__global__ void __launch_bounds__(HR_TILE)
head_reparam_fwd_kernel(const float* __restrict__ h, const float* __restrict__ Wm, const float* __restrict__ bm, const float* __restrict__ Ws,
                        const float* __restrict__ bs,
                        const float* __restrict__ eps, float* __restrict__ mu, float* __restrict__ sigma, float* __restrict__ z,
                        float* __restrict__ angles, float* __restrict__ log_q, int64_t total, int64_t B, int Din, int krt) {
    constexpr int DM = hr_mean_rows(MODE), DT = DM + 3;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                        // [(Din+1)][12]
    float* s_e = s_w + (Din + 1) * HR_WPAD;                   // eps in, Euler angles out
    float* s_z = s_e + HR_TILE * 3;                           // z out
    float* s_h = s_z + HR_TILE * 9;                           // [256][Din]
    const int64_t i0 = int64_t(blockIdx.x) * HR_TILE;
    const int rows = int(min(int64_t(HR_TILE), total - i0));
    hr_stage_h(s_h, h, i0, rows, B, Din);
    tile_g2s(s_e, eps + i0 * 3, rows * 3);
    hr_stage_w(s_w, Wm, bm, Ws, bs, DM, Din);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    if (t < rows) {
        float pre[9], m[9], sg[3], ep[3], zr[9], e[3], lq;
        hr_linear<DT>(s_w, s_h + t * Din, Din, pre);
        hr_mean_fwd<MODE>(pre, m);
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = hr_softplus(pre[DM + j]); ep[j] = s_e[t * 3 + j]; }
        reparam_sample_fwd<float, KT, EULER>(m, sg, ep, krt, log_q != nullptr, zr, e, &lq);
#pragma unroll
        for (int j = 0; j < 9; ++j) s_z[t * 9 + j] = zr[j];
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = e[j];
        }
        if (log_q != nullptr) log_q[i0 + t] = lq;
        if (i0 + t < B) {                                     // module attributes mu_lie / sigma: first sample set only
            if (mu != nullptr) {
#pragma unroll
                for (int j = 0; j < 9; ++j) mu[(i0 + t) * 9 + j] = m[j];
            }
            if (sigma != nullptr) {
#pragma unroll
                for (int j = 0; j < 3; ++j) sigma[(i0 + t) * 3 + j] = sg[j];
            }
        }
    }
    __syncthreads();
    if (z != nullptr) tile_s2g(z + i0 * 9, s_z, rows * 9);
    if (EULER) tile_s2g(angles + i0 * 3, s_e, rows * 3);
}

// ------------------------------------------------------------------ backward
// g_h (total, Din) per sample (summed over n by the caller); partial [gridDim.x][DT * (Din+1)]: the CTA's share of
// [g_W | g_bias] (row j: Din weight gradients then the bias gradient).
template <int MODE, int KT, bool EULER>
__global__ void __launch_bounds__(HR_TILE)
head_reparam_bwd_kernel(const float* __restrict__ h, const float* __restrict__ Wm, const float* __restrict__ bm, const float* __restrict__ Ws,
                        const float* __restrict__ bs,
                        const float* __restrict__ eps, const float* __restrict__ gz, const float* __restrict__ gangles,
                        const float* __restrict__ glq, const float* __restrict__ gmu_ext, const float* __restrict__ gsg_ext,
                        float* __restrict__ gh, float* __restrict__ partial, int64_t total, int64_t B, int Din, int krt) {
    constexpr int DM = hr_mean_rows(MODE), DT = DM + 3;
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                        // [(Din+1)][12]
    float* s_e = s_w + (Din + 1) * HR_WPAD;                   // eps
    float* s_g = s_e + HR_TILE * 3;                           // gz in
    float* s_a = s_g + HR_TILE * 9;                           // g_angles in
    float* s_p = s_a + HR_TILE * 3;                           // [256][12] g_pre
    float* s_h = s_p + HR_TILE * HR_WPAD;                     // [256][Din] h in
    float* s_o = s_h + HR_TILE * Din;                         // [256][Din] g_h out
    const int64_t i0 = int64_t(blockIdx.x) * HR_TILE;
    const int rows = int(min(int64_t(HR_TILE), total - i0));
    hr_stage_h(s_h, h, i0, rows, B, Din);
    tile_g2s(s_e, eps + i0 * 3, rows * 3);
    if (gz != nullptr) tile_g2s(s_g, gz + i0 * 9, rows * 9);
    if (EULER) tile_g2s(s_a, gangles + i0 * 3, rows * 3);
    hr_stage_w(s_w, Wm, bm, Ws, bs, DM, Din);
    tile_async_wait();
    __syncthreads();
    const int t = threadIdx.x;
    float gpre[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) gpre[j] = 0.f;
    if (t < rows) {
        float pre[9], m[9], sg[3], ep[3], G[9], ge[3] = {0.f, 0.f, 0.f}, gm[9], gsg[3];
        hr_linear<DT>(s_w, s_h + t * Din, Din, pre);
        hr_mean_fwd<MODE>(pre, m);
#pragma unroll
        for (int j = 0; j < 3; ++j) { sg[j] = hr_softplus(pre[DM + j]); ep[j] = s_e[t * 3 + j]; }
#pragma unroll
        for (int j = 0; j < 9; ++j) G[j] = gz != nullptr ? s_g[t * 9 + j] : 0.f;
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ge[j] = s_a[t * 3 + j];
        }
        const float gl = glq != nullptr ? glq[i0 + t] : 0.f;
        reparam_sample_bwd<float, KT, EULER>(m, sg, ep, G, ge, gl, glq != nullptr, krt, gm, gsg);
        // gradients that reach mu (B,3,3) / sigma (B,3) from outside the sampler (the modules expose them as mu_lie / sigma:
        // kl(), regularisers): mu and sigma are broadcast over n, so each is added once, in the datapoint's first sample
        if (i0 + t < B) {
            if (gmu_ext != nullptr) {
#pragma unroll
                for (int j = 0; j < 9; ++j) gm[j] += __ldg(gmu_ext + (i0 + t) * 9 + j);
            }
            if (gsg_ext != nullptr) {
#pragma unroll
                for (int j = 0; j < 3; ++j) gsg[j] += __ldg(gsg_ext + (i0 + t) * 3 + j);
            }
        }
        hr_mean_bwd<MODE>(pre, gm, gpre);
#pragma unroll
        for (int j = 0; j < 3; ++j) gpre[DM + j] = gsg[j] * hr_softplus_grad(pre[DM + j]);
        // g_h = W^T g_pre
        const float4* w4 = reinterpret_cast<const float4*>(s_w);
        for (int d = 0; d < Din; ++d) {
            const float4 a = w4[d * 3], b = w4[d * 3 + 1], c = w4[d * 3 + 2];
            float acc = a.x * gpre[0];
            acc = fmaf(a.y, gpre[1], acc); acc = fmaf(a.z, gpre[2], acc); acc = fmaf(a.w, gpre[3], acc);
            acc = fmaf(b.x, gpre[4], acc); acc = fmaf(b.y, gpre[5], acc); acc = fmaf(b.z, gpre[6], acc); acc = fmaf(b.w, gpre[7], acc);
            acc = fmaf(c.x, gpre[8], acc);
            s_o[t * Din + d] = acc;
        }
    }
    {
        float4* p4 = reinterpret_cast<float4*>(s_p + t * HR_WPAD);          // rows past the end contribute zeros
        p4[0] = make_float4(gpre[0], gpre[1], gpre[2], gpre[3]);
        p4[1] = make_float4(gpre[4], gpre[5], gpre[6], gpre[7]);
        p4[2] = make_float4(gpre[8], 0.f, 0.f, 0.f);
    }
    __syncthreads();
    tile_s2g(gh + i0 * Din, s_o, rows * Din);
    // [g_W | g_b] of this CTA: output o = j * (Din+1) + d  (d == Din: bias), summed over the tile's rows in row order
    const int NO = DT * (Din + 1);
    for (int o = t; o < NO; o += HR_TILE) {
        const int j = o / (Din + 1), d = o - j * (Din + 1);
        float acc = 0.f;
        if (d < Din) {
            for (int r = 0; r < rows; ++r) acc = fmaf(s_p[r * HR_WPAD + j], s_h[r * Din + d], acc);
        } else {
            for (int r = 0; r < rows; ++r) acc += s_p[r * HR_WPAD + j];
        }
        partial[int64_t(blockIdx.x) * NO + o] = acc;
    }
}

// partial [nblk][NO] -> out[NO], fixed order
__global__ void __launch_bounds__(256)
head_reduce_partials(const float* __restrict__ partial, float* __restrict__ out, int64_t nblk, int NO) {
    __shared__ float red[8][33];
    const int o = blockIdx.x * 32 + threadIdx.x;
    float acc = 0.f;
    if (o < NO)
        for (int64_t b = threadIdx.y; b < nblk; b += 8) acc += partial[b * NO + o];
    red[threadIdx.y][threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.y == 0 && o < NO) {
        float a = 0.f;
#pragma unroll
        for (int y = 0; y < 8; ++y) a += red[y][threadIdx.x];
        out[o] = a;
    }
}

static int hr_check(const char* name, int64_t n, int64_t B, int Din, int mode, int k) {
    if (n < 0 || B < 0 || k < 0 || Din <= 0) { set_error("%s: bad sizes", name); return LV_ERR_ARG; }
    if (mode < HR_ALG || mode > HR_S2S1) { set_error("%s: unknown mean mode %d", name, mode); return LV_ERR_ARG; }
    if (Din > HR_MAX_DIN) { set_error("%s: more than %d input features unsupported", name, HR_MAX_DIN); return LV_ERR_UNSUPPORTED; }
    if (k > 64) { set_error("%s: k=%d winding terms unsupported (max 64)", name, k); return LV_ERR_UNSUPPORTED; }
    if ((n * B + HR_TILE - 1) / HR_TILE > 0x7fffffffLL) { set_error("%s: too many samples", name); return LV_ERR_ARG; }
    return LV_OK;
}

template <int MODE, bool EULER>
static int hr_launch_fwd(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs, const float* eps, float* mu, float* sigma, float* z,
                         float* angles, float* log_q, int64_t total, int64_t B, int Din, int k, cudaStream_t st) {
    const unsigned grid = unsigned((total + HR_TILE - 1) / HR_TILE);
    const size_t smem = size_t((Din + 1) * HR_WPAD + HR_TILE * (3 + 9 + Din)) * 4;
    if (k == 3) head_reparam_fwd_kernel<MODE, 3, EULER><<<grid, HR_TILE, smem, st>>>(h, Wm, bm, Ws, bs, eps, mu, sigma, z, angles, log_q, total, B, Din, k);
    else if (k == 10) head_reparam_fwd_kernel<MODE, 10, EULER><<<grid, HR_TILE, smem, st>>>(h, Wm, bm, Ws, bs, eps, mu, sigma, z, angles, log_q, total, B, Din, k);
    else head_reparam_fwd_kernel<MODE, 0, EULER><<<grid, HR_TILE, smem, st>>>(h, Wm, bm, Ws, bs, eps, mu, sigma, z, angles, log_q, total, B, Din, k);
    return check_launch("so3_head_reparam_fwd");
}
template <int MODE, bool EULER, int KT>
static int hr_launch_bwd_k(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs, const float* eps, const float* gz, const float* gangles,
                           const float* glq, const float* gmu_ext, const float* gsg_ext, float* gh, float* partial, int64_t total, int64_t B, int Din, int k,
                           cudaStream_t st) {
    const unsigned grid = unsigned((total + HR_TILE - 1) / HR_TILE);
    const size_t smem = size_t((Din + 1) * HR_WPAD + HR_TILE * (3 + 9 + 3 + HR_WPAD + 2 * Din)) * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(head_reparam_bwd_kernel<MODE, KT, EULER>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e != cudaSuccess) { set_error("so3_head_reparam_bwd: cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e)); return int(e); }
    }
    head_reparam_bwd_kernel<MODE, KT, EULER><<<grid, HR_TILE, smem, st>>>(h, Wm, bm, Ws, bs, eps, gz, gangles, glq, gmu_ext, gsg_ext, gh, partial, total, B, Din, k);
    return check_launch("so3_head_reparam_bwd");
}
template <int MODE, bool EULER>
static int hr_launch_bwd(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs, const float* eps, const float* gz, const float* gangles,
                         const float* glq, const float* gmu_ext, const float* gsg_ext, float* gh, float* partial, int64_t total, int64_t B, int Din, int k,
                         cudaStream_t st) {
    if (k == 3) return hr_launch_bwd_k<MODE, EULER, 3>(h, Wm, bm, Ws, bs, eps, gz, gangles, glq, gmu_ext, gsg_ext, gh, partial, total, B, Din, k, st);
    if (k == 10) return hr_launch_bwd_k<MODE, EULER, 10>(h, Wm, bm, Ws, bs, eps, gz, gangles, glq, gmu_ext, gsg_ext, gh, partial, total, B, Din, k, st);
    return hr_launch_bwd_k<MODE, EULER, 0>(h, Wm, bm, Ws, bs, eps, gz, gangles, glq, gmu_ext, gsg_ext, gh, partial, total, B, Din, k, st);
}

}  // namespace lv

#define HR_DISPATCH(mode, euler, FN, ...)                                                          \
    ((mode) == lv::HR_ALG ? ((euler) ? lv::FN<lv::HR_ALG, true>(__VA_ARGS__) : lv::FN<lv::HR_ALG, false>(__VA_ARGS__))       \
     : (mode) == lv::HR_QUAT ? ((euler) ? lv::FN<lv::HR_QUAT, true>(__VA_ARGS__) : lv::FN<lv::HR_QUAT, false>(__VA_ARGS__)) \
     : (mode) == lv::HR_S2S1 ? ((euler) ? lv::FN<lv::HR_S2S1, true>(__VA_ARGS__) : lv::FN<lv::HR_S2S1, false>(__VA_ARGS__)) \
                             : ((euler) ? lv::FN<lv::HR_S2S2, true>(__VA_ARGS__) : lv::FN<lv::HR_S2S2, false>(__VA_ARGS__)))

// ====================================================================== C ABI
extern "C" int64_t lv_so3_head_reparam_bwd_workspace_floats(int64_t n, int64_t B, int Din, int mode) {
    if (n < 0 || B < 0 || Din <= 0 || mode < lv::HR_ALG || mode > lv::HR_S2S1) return -1;
    return ((n * B + lv::HR_TILE - 1) / lv::HR_TILE) * int64_t(lv::hr_mean_rows(mode) + 3) * (Din + 1);
}

extern "C" int lv_so3_head_reparam_fwd_f32(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs, const float* eps, float* mu,
                                           float* sigma, float* z, float* angles, float* log_q, int64_t n, int64_t B, int Din,
                                           int mode, int k, void* stream) {
    int rc = lv::hr_check("so3_head_reparam_fwd", n, B, Din, mode, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!h || !Wm || !bm || !Ws || !bs || !eps || (!z && !angles)) { lv::set_error("so3_head_reparam_fwd: null pointer"); return LV_ERR_ARG; }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return HR_DISPATCH(mode, angles != nullptr, hr_launch_fwd, h, Wm, bm, Ws, bs, eps, mu, sigma, z, angles, log_q, total, B, Din, k, st);
}

extern "C" int lv_so3_head_reparam_bwd_f32(const float* h, const float* Wm, const float* bm, const float* Ws, const float* bs, const float* eps, const float* gz,
                                           const float* gangles, const float* glq, const float* gmu, const float* gsigma, float* gh, float* gWb,
                                           float* workspace, int64_t workspace_floats, int64_t n, int64_t B, int Din, int mode, int k, void* stream) {
    int rc = lv::hr_check("so3_head_reparam_bwd", n, B, Din, mode, k);
    if (rc) return rc;
    const int64_t total = n * B;
    const int NO = (lv::hr_mean_rows(mode) + 3) * (Din + 1);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (!gWb) { lv::set_error("so3_head_reparam_bwd: null pointer"); return LV_ERR_ARG; }
    if (total == 0) {
        cudaError_t e = cudaMemsetAsync(gWb, 0, size_t(NO) * 4, st);
        if (e != cudaSuccess) { lv::set_error("so3_head_reparam_bwd: memset: %s", cudaGetErrorString(e)); return int(e); }
        return LV_OK;
    }
    if (!h || !Wm || !bm || !Ws || !bs || !eps || !gh || (!gz && !gangles && !glq && !gmu && !gsigma)) { lv::set_error("so3_head_reparam_bwd: null pointer"); return LV_ERR_ARG; }
    const int64_t need = lv_so3_head_reparam_bwd_workspace_floats(n, B, Din, mode);
    if (!workspace || workspace_floats < need) { lv::set_error("so3_head_reparam_bwd: workspace of %lld floats required", (long long)need); return LV_ERR_ARG; }
    rc = HR_DISPATCH(mode, gangles != nullptr, hr_launch_bwd, h, Wm, bm, Ws, bs, eps, gz, gangles, glq, gmu, gsigma, gh, workspace, total, B, Din, k, st);
    if (rc) return rc;
    const int64_t nblk = (total + lv::HR_TILE - 1) / lv::HR_TILE;
    lv::head_reduce_partials<<<(NO + 31) / 32, dim3(32, 8), 0, st>>>(workspace, gWb, nblk, NO);
    return lv::check_launch("head_reduce_partials");
}
