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
// per sample; each CTA stages its 256-sample tile through shared memory (rows of 9 and 3
// floats are odd strides -> conflict-free LDS), all arithmetic in registers, outputs written
// back through the same staging buffers.  A tile is one contiguous span of every tensor, so
// full tiles move with TMA bulk copies issued by ONE thread (cp.async.bulk + mbarrier in,
// cp.async.bulk.global.shared::cta out): the kernels are issue-bound and the per-thread
// cp.async / store loops (addresses, alignment and bound checks: ~120 + ~60 of the ~800
// instructions a thread executed) are what this removes.  Ragged last tiles, tiles that wrap
// over the broadcast rows and unaligned tensors take the cp.async path.
// HBM-bound: 100 B/sample forward, 148 B/sample backward.
#include "common.cuh"
#include "reparam_core.cuh"

namespace lv {

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

// A full tile can move with TMA bulk copies when every tensor is 16-byte aligned (host flag), the tile does not wrap
// over the broadcast mu / sigma rows, and its first broadcast row starts on a 16-byte boundary (rows are 36 / 12 bytes
// in float, 72 / 24 in double: b0 a multiple of 4 covers all of them; the tile's own start i0 is a multiple of TILE).
template <int TILE>
__device__ __forceinline__ bool bulk_tile_ok(int aligned16, bool full, int64_t b0, int64_t B) {
    return aligned16 != 0 && full && b0 + TILE <= B && (b0 & 3) == 0;
}

// ------------------------------------------------------------------ forward
// EULER: additionally emit the ZYZ Euler angles of z (group_matrix_to_eazyz, lie_tools.py:178-180 -- what
// VAE.decode feeds the action decoder, vae.py:182) from the registers that hold z; z itself is then optional.
template <typename T, int KT, bool EULER, int TILE>
__global__ void __launch_bounds__(TILE)
so3_reparam_fwd_kernel(const T* __restrict__ mu, const T* __restrict__ sigma, const T* __restrict__ eps,
                       T* __restrict__ z, T* __restrict__ angles, T* __restrict__ log_q, int64_t total,
                       int64_t B, int krt, int aligned16, uint64_t seed, uint64_t offset) {
    __shared__ __align__(16) T s_m[TILE * 9];   // mu in, z out (same row, same thread)
    __shared__ __align__(16) T s_s[TILE * 3];
    __shared__ __align__(16) T s_e[TILE * 3];   // eps in, Euler angles out
    __shared__ __align__(8) uint64_t s_bar;
    const int64_t i0 = int64_t(blockIdx.x) * TILE;
    const int rows = int(min(int64_t(TILE), total - i0));
    const bool full = rows == TILE;       // every CTA but the last: compile-time copy loops
    const int64_t b0 = i0 < B ? i0 : i0 % B;
    const bool bulk = bulk_tile_ok<TILE>(aligned16, full, b0, B);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&s_bar, uint32_t(TILE * (eps != nullptr ? 15 : 12) * sizeof(T)));
            tma_load(s_m, mu + b0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_bar);
            tma_load(s_s, sigma + b0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_bar);
            if (eps != nullptr) tma_load(s_e, eps + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_bar);
        }
        __syncthreads();                  // the barrier word is initialised before anyone polls it
        mbar_wait(&s_bar, 0);
    } else {
        stage_bcast<T, 9, TILE>(s_m, mu, i0, rows, B);
        stage_bcast<T, 3, TILE>(s_s, sigma, i0, rows, B);
        if (eps != nullptr) {
            if (full) tile_g2s_full<T, TILE * 3, TILE>(s_e, eps + i0 * 3);
            else tile_g2s(s_e, eps + i0 * 3, rows * 3);
        }
        tile_async_wait();
        __syncthreads();
    }
    const int t = threadIdx.x;
    if (t < rows) {
        T m[9], sg[3], ep[3], zr[9], e[3], lq;
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = s_m[t * 9 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) sg[j] = s_s[t * 3 + j];
        if (eps != nullptr) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ep[j] = s_e[t * 3 + j];
        } else {
            philox_normal3<T>(seed, offset + uint64_t(i0 + t), ep);
        }
        reparam_sample_fwd<T, KT, EULER>(m, sg, ep, krt, log_q != nullptr, zr, e, &lq);
#pragma unroll
        for (int j = 0; j < 9; ++j) s_m[t * 9 + j] = zr[j];
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = e[j];
        }
        if (log_q != nullptr) log_q[i0 + t] = lq;
    }
    if (bulk) {
        fence_proxy_async_smem();         // this thread's smem writes -> visible to the copy engine
        __syncthreads();
        if (threadIdx.x == 0) {
            if (z != nullptr) tma_store(z + i0 * 9, s_m, uint32_t(TILE * 9 * sizeof(T)));
            if (EULER) tma_store(angles + i0 * 3, s_e, uint32_t(TILE * 3 * sizeof(T)));
            tma_store_commit_wait();
        }
        return;
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
                       T* __restrict__ gmu, T* __restrict__ gsigma, int64_t total, int64_t B, int krt, int aligned16,
                       uint64_t seed, uint64_t offset) {
    __shared__ __align__(16) T s_m[TILE * 9];
    __shared__ __align__(16) T s_g[TILE * 9];   // gz in, g_mu out
    __shared__ __align__(16) T s_s[TILE * 3];
    __shared__ __align__(16) T s_e[TILE * 3];   // eps in, g_sigma out
    __shared__ __align__(16) T s_a[EULER ? TILE * 3 : 4];   // g_angles in
    __shared__ __align__(8) uint64_t s_bar;
    const int64_t i0 = int64_t(blockIdx.x) * TILE;
    const int rows = int(min(int64_t(TILE), total - i0));
    const bool full = rows == TILE;
    const int64_t b0 = i0 < B ? i0 : i0 % B;
    const bool bulk = bulk_tile_ok<TILE>(aligned16, full, b0, B);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            mbar_expect_tx(&s_bar, uint32_t(TILE * (12 + (eps != nullptr ? 3 : 0) + (gz != nullptr ? 9 : 0) + (EULER ? 3 : 0)) * sizeof(T)));
            tma_load(s_m, mu + b0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_bar);
            tma_load(s_s, sigma + b0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_bar);
            if (eps != nullptr) tma_load(s_e, eps + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_bar);
            if (gz != nullptr) tma_load(s_g, gz + i0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_bar);
            if (EULER) tma_load(s_a, gangles + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_bar);
        }
        __syncthreads();
        mbar_wait(&s_bar, 0);
    } else {
        stage_bcast<T, 9, TILE>(s_m, mu, i0, rows, B);
        stage_bcast<T, 3, TILE>(s_s, sigma, i0, rows, B);
        if (full) {
            if (eps != nullptr) tile_g2s_full<T, TILE * 3, TILE>(s_e, eps + i0 * 3);
            if (gz != nullptr) tile_g2s_full<T, TILE * 9, TILE>(s_g, gz + i0 * 9);
            if (EULER) tile_g2s_full<T, TILE * 3, TILE>(s_a, gangles + i0 * 3);
        } else {
            if (eps != nullptr) tile_g2s(s_e, eps + i0 * 3, rows * 3);
            if (gz != nullptr) tile_g2s(s_g, gz + i0 * 9, rows * 9);
            if (EULER) tile_g2s(s_a, gangles + i0 * 3, rows * 3);
        }
        tile_async_wait();
        __syncthreads();
    }
    const int t = threadIdx.x;
    if (t < rows) {
        T m[9], G[9], sg[3], ep[3], ge[3] = {T(0), T(0), T(0)}, gm[9], gsg[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) { m[j] = s_m[t * 9 + j]; G[j] = gz != nullptr ? s_g[t * 9 + j] : T(0); }
#pragma unroll
        for (int j = 0; j < 3; ++j) sg[j] = s_s[t * 3 + j];
        if (eps != nullptr) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ep[j] = s_e[t * 3 + j];
        } else {
            philox_normal3<T>(seed, offset + uint64_t(i0 + t), ep);
        }
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ge[j] = s_a[t * 3 + j];
        }
        const T gl = glq != nullptr ? glq[i0 + t] : T(0);
        reparam_sample_bwd<T, KT, EULER>(m, sg, ep, G, ge, gl, glq != nullptr, krt, gm, gsg);
#pragma unroll
        for (int j = 0; j < 9; ++j) s_g[t * 9 + j] = gm[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) s_e[t * 3 + j] = gsg[j];
    }
    if (bulk) {
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            tma_store(gmu + i0 * 9, s_g, uint32_t(TILE * 9 * sizeof(T)));
            tma_store(gsigma + i0 * 3, s_e, uint32_t(TILE * 3 * sizeof(T)));
            tma_store_commit_wait();
        }
        return;
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

// ------------------------------------------------------------------ persistent, double-buffered variants
// Used when every tile can move with TMA bulk copies (16-byte aligned tensors; n == 1, or B a multiple of the tile so
// that no tile wraps over the broadcast rows) and there are enough full tiles to keep every SM's resident CTAs busy for
// several rounds.  A CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... through TWO sets of staging buffers:
// thread 0 issues the bulk loads of the next tile before the CTA starts on the current one, so the load latency the
// one-tile-per-CTA kernels expose at their barrier (ncu: 27 % of the stall samples) overlaps the arithmetic.
//   iteration it, stage s = it % NS:  [thread 0] wait until the store issued from stage (it+1) % NS has been read out
//                                                (NS = 2: the previous iteration's; NS = 3: the one before, a full
//                                                iteration of slack), then load(tile it+1 -> stage (it+1) % NS)
//                                     wait full[s] (parity (it / NS) & 1)
//                                     compute in place; fence.proxy.async; __syncthreads
//                                     [thread 0] bulk store(stage s), commit
// Forward: three stages (45 KB) -- with two, thread 0 waits for its store right after issuing it and the other 255
// threads wait for thread 0 at the next barrier; backward: two stages (55 KB; a third would cost a resident CTA).
// The ragged tail (total % TILE samples) is a separate launch of the kernels above.
template <typename T, int TILE, bool EULER> struct RpFwdStage {
    T m[TILE * 9];      // mu in, z out
    T s[TILE * 3];
    T e[TILE * 3];      // eps in, Euler angles out
};
template <typename T, int TILE, bool EULER> struct RpBwdStage {
    T m[TILE * 9];
    T g[TILE * 9];      // gz in, g_mu out
    T s[TILE * 3];
    T e[TILE * 3];      // eps in, g_sigma out
    T a[EULER ? TILE * 3 : 4];   // g_angles in
};

template <typename T, int KT, bool EULER, int TILE, int NS>
__global__ void __launch_bounds__(TILE)
so3_reparam_fwd_pipe_kernel(const T* __restrict__ mu, const T* __restrict__ sigma, const T* __restrict__ eps,
                            T* __restrict__ z, T* __restrict__ angles, T* __restrict__ log_q, int64_t ntiles,
                            int64_t B, int krt, uint64_t seed, uint64_t offset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Stage = RpFwdStage<T, TILE, EULER>;
    Stage* st = reinterpret_cast<Stage*>(smem_raw);
    __shared__ __align__(8) uint64_t s_full[NS];
    const int t = threadIdx.x;
    auto load = [&](int64_t tile, int s) {
        const int64_t i0 = tile * TILE, b0 = i0 < B ? i0 : i0 % B;
        mbar_expect_tx(&s_full[s], uint32_t(TILE * (eps != nullptr ? 15 : 12) * sizeof(T)));
        tma_load(st[s].m, mu + b0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_full[s]);
        tma_load(st[s].s, sigma + b0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_full[s]);
        if (eps != nullptr) tma_load(st[s].e, eps + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_full[s]);
    };
    if (t == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) mbar_init(&s_full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (int64_t(blockIdx.x) < ntiles) load(blockIdx.x, 0);
    }
    __syncthreads();
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % NS;
        if (t == 0 && tile + gridDim.x < ntiles) {
            tma_store_wait_read<NS - 2>();
            load(tile + gridDim.x, (it + 1) % NS);
        }
        mbar_wait(&s_full[s], uint32_t(it / NS) & 1u);
        const int64_t i0 = tile * TILE;
        T m[9], sg[3], ep[3], zr[9], e[3], lq;
#pragma unroll
        for (int j = 0; j < 9; ++j) m[j] = st[s].m[t * 9 + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) sg[j] = st[s].s[t * 3 + j];
        if (eps != nullptr) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ep[j] = st[s].e[t * 3 + j];
        } else {
            philox_normal3<T>(seed, offset + uint64_t(i0 + t), ep);
        }
        reparam_sample_fwd<T, KT, EULER>(m, sg, ep, krt, log_q != nullptr, zr, e, &lq);
#pragma unroll
        for (int j = 0; j < 9; ++j) st[s].m[t * 9 + j] = zr[j];
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) st[s].e[t * 3 + j] = e[j];
        }
        if (log_q != nullptr) log_q[i0 + t] = lq;
        fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            if (z != nullptr) tma_store(z + i0 * 9, st[s].m, uint32_t(TILE * 9 * sizeof(T)));
            if (EULER) tma_store(angles + i0 * 3, st[s].e, uint32_t(TILE * 3 * sizeof(T)));
            tma_store_commit();
        }
    }
    if (t == 0) tma_store_wait_read<0>();      // the staging buffers stay valid until the copy engine has read them
}

template <typename T, int KT, bool EULER, int TILE, int NS>
__global__ void __launch_bounds__(TILE)
so3_reparam_bwd_pipe_kernel(const T* __restrict__ mu, const T* __restrict__ sigma, const T* __restrict__ eps,
                            const T* __restrict__ gz, const T* __restrict__ gangles, const T* __restrict__ glq,
                            T* __restrict__ gmu, T* __restrict__ gsigma, int64_t ntiles, int64_t B, int krt,
                            uint64_t seed, uint64_t offset) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Stage = RpBwdStage<T, TILE, EULER>;
    Stage* st = reinterpret_cast<Stage*>(smem_raw);
    __shared__ __align__(8) uint64_t s_full[NS];
    const int t = threadIdx.x;
    auto load = [&](int64_t tile, int s) {
        const int64_t i0 = tile * TILE, b0 = i0 < B ? i0 : i0 % B;
        mbar_expect_tx(&s_full[s], uint32_t(TILE * (12 + (eps != nullptr ? 3 : 0) + (gz != nullptr ? 9 : 0) + (EULER ? 3 : 0)) * sizeof(T)));
        tma_load(st[s].m, mu + b0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_full[s]);
        tma_load(st[s].s, sigma + b0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_full[s]);
        if (eps != nullptr) tma_load(st[s].e, eps + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_full[s]);
        if (gz != nullptr) tma_load(st[s].g, gz + i0 * 9, uint32_t(TILE * 9 * sizeof(T)), &s_full[s]);
        if (EULER) tma_load(st[s].a, gangles + i0 * 3, uint32_t(TILE * 3 * sizeof(T)), &s_full[s]);
    };
    if (t == 0) {
#pragma unroll
        for (int q = 0; q < NS; ++q) mbar_init(&s_full[q], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (int64_t(blockIdx.x) < ntiles) load(blockIdx.x, 0);
    }
    __syncthreads();
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int s = it % NS;
        if (t == 0 && tile + gridDim.x < ntiles) {
            tma_store_wait_read<NS - 2>();
            load(tile + gridDim.x, (it + 1) % NS);
        }
        const int64_t i0 = tile * TILE;
        const T gl = glq != nullptr ? glq[i0 + t] : T(0);
        mbar_wait(&s_full[s], uint32_t(it / NS) & 1u);
        T m[9], G[9], sg[3], ep[3], ge[3] = {T(0), T(0), T(0)}, gm[9], gsg[3];
#pragma unroll
        for (int j = 0; j < 9; ++j) { m[j] = st[s].m[t * 9 + j]; G[j] = gz != nullptr ? st[s].g[t * 9 + j] : T(0); }
#pragma unroll
        for (int j = 0; j < 3; ++j) sg[j] = st[s].s[t * 3 + j];
        if (eps != nullptr) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ep[j] = st[s].e[t * 3 + j];
        } else {
            philox_normal3<T>(seed, offset + uint64_t(i0 + t), ep);
        }
        if (EULER) {
#pragma unroll
            for (int j = 0; j < 3; ++j) ge[j] = st[s].a[t * 3 + j];
        }
        reparam_sample_bwd<T, KT, EULER>(m, sg, ep, G, ge, gl, glq != nullptr, krt, gm, gsg);
#pragma unroll
        for (int j = 0; j < 9; ++j) st[s].g[t * 9 + j] = gm[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) st[s].e[t * 3 + j] = gsg[j];
        fence_proxy_async_smem();
        __syncthreads();
        if (t == 0) {
            tma_store(gmu + i0 * 9, st[s].g, uint32_t(TILE * 9 * sizeof(T)));
            tma_store(gsigma + i0 * 3, st[s].e, uint32_t(TILE * 3 * sizeof(T)));
            tma_store_commit();
        }
    }
    if (t == 0) tma_store_wait_read<0>();      // the staging buffers stay valid until the copy engine has read them
}

// eps (rows, 3) of the Philox stream the fused kernels draw from when they are given no eps: row i = sample offset + i
template <typename T>
__global__ void __launch_bounds__(256)
philox_normal_kernel(T* __restrict__ out, int64_t rows, uint64_t seed, uint64_t offset) {
    const int64_t i = int64_t(blockIdx.x) * 256 + threadIdx.x;
    if (i >= rows) return;
    T ep[3];
    philox_normal3<T>(seed, offset + uint64_t(i), ep);
    out[i * 3] = ep[0]; out[i * 3 + 1] = ep[1]; out[i * 3 + 2] = ep[2];
}

}  // namespace lv

// ====================================================================== C ABI
static int reparam_check(const char* name, int64_t n, int64_t B, int k) {
    if (n < 0 || B < 0 || k < 0) { lv::set_error("%s: negative size", name); return LV_ERR_ARG; }
    if (k > 64) { lv::set_error("%s: k=%d winding terms unsupported (max 64)", name, k); return LV_ERR_UNSUPPORTED; }
    if ((n * B + lv::RP_TILE / 2 - 1) / (lv::RP_TILE / 2) > 0x7fffffffLL) { lv::set_error("%s: too many samples", name); return LV_ERR_ARG; }
    return LV_OK;
}

template <typename... P>
static int aligned16(const P*... p) {
    return (((reinterpret_cast<uintptr_t>(p)) | ...) & 15u) == 0;
}

// persistent launch geometry: the CTAs that are resident at once, trimmed so that every CTA walks the same number of
// tiles (a partial last round would leave most SMs idle for one tile time); 0 = use the one-tile-per-CTA kernels.
// `cap_cache` (one array per kernel instantiation, indexed by device) keeps SMs x resident CTAs after the first launch on
// a device: the occupancy query and the shared-memory opt-in cost ~10 us of host time, as much as a small launch.
constexpr int RP_MAX_DEVICES = 64;
template <typename K>
static int pipe_grid(K kernel, int threads, size_t smem, int64_t ntiles, int* cap_cache, int* grid) {
    *grid = 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    int64_t cap = (e == cudaSuccess && dev >= 0 && dev < RP_MAX_DEVICES) ? cap_cache[dev] : 0;
    if (cap == 0) {
        int sms = 0, per_sm = 0;
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e == cudaSuccess && smem > 48 * 1024) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
        if (e != cudaSuccess) { lv::set_error("so3_reparam: launch geometry query failed (%s)", cudaGetErrorString(e)); return int(e); }
        cap = int64_t(sms) * per_sm;
        if (cap > 0 && dev >= 0 && dev < RP_MAX_DEVICES) cap_cache[dev] = int(cap);
    }
    if (cap < 1 || ntiles < 2 * cap) return LV_OK;          // too few tiles to pipeline: one tile per CTA
    const int64_t rounds = (ntiles + cap - 1) / cap;
    *grid = int((ntiles + rounds - 1) / rounds);
    return LV_OK;
}

template <typename T, int KT, bool EULER, int TILE>
static int reparam_fwd_launch(const T* mu, const T* sigma, const T* eps, T* z, T* angles, T* log_q, int64_t n, int64_t B, int k,
                              cudaStream_t st, uint64_t seed, uint64_t offset) {
    const int64_t total = n * B;
    const int al = aligned16(mu, sigma, eps, z, angles);      // null pointers count as aligned
    int64_t done = 0;
    if (al && (n == 1 || B % TILE == 0)) {
        constexpr int NS = 3;
        constexpr size_t SMEM = NS * sizeof(lv::RpFwdStage<T, TILE, EULER>);
        const int64_t nfull = total / TILE;
        static int cap_cache[RP_MAX_DEVICES] = {0};
        int grid = 0;
        int rc = pipe_grid(lv::so3_reparam_fwd_pipe_kernel<T, KT, EULER, TILE, NS>, TILE, SMEM, nfull, cap_cache, &grid);
        if (rc) return rc;
        if (grid > 0) {
            lv::so3_reparam_fwd_pipe_kernel<T, KT, EULER, TILE, NS><<<grid, TILE, SMEM, st>>>(mu, sigma, eps, z, angles, log_q, nfull, B, k, seed, offset);
            done = nfull * TILE;                              // the ragged tail (n == 1 only) follows below
        }
    }
    if (done < total) {
        const int64_t rest = total - done;                    // done > 0 implies n == 1: the tail is its own (B = rest) problem
        const unsigned grid = unsigned((rest + TILE - 1) / TILE);
        lv::so3_reparam_fwd_kernel<T, KT, EULER, TILE><<<grid, TILE, 0, st>>>(
            mu + (done ? done * 9 : 0), sigma + (done ? done * 3 : 0), eps ? eps + done * 3 : nullptr, z ? z + done * 9 : nullptr,
            angles ? angles + done * 3 : nullptr, log_q ? log_q + done : nullptr, rest, done ? rest : B, k, al, seed, offset + uint64_t(done));
    }
    return LV_OK;
}

template <typename T, int KT, bool EULER, int TILE>
static int reparam_bwd_launch(const T* mu, const T* sigma, const T* eps, const T* gz, const T* gangles, const T* glq, T* gmu,
                              T* gsigma, int64_t n, int64_t B, int k, cudaStream_t st, uint64_t seed, uint64_t offset) {
    const int64_t total = n * B;
    const int al = aligned16(mu, sigma, eps, gz, gangles, gmu, gsigma);
    int64_t done = 0;
    if (al && (n == 1 || B % TILE == 0)) {
        constexpr int NS = 2;
        constexpr size_t SMEM = NS * sizeof(lv::RpBwdStage<T, TILE, EULER>);
        const int64_t nfull = total / TILE;
        static int cap_cache[RP_MAX_DEVICES] = {0};
        int grid = 0;
        int rc = pipe_grid(lv::so3_reparam_bwd_pipe_kernel<T, KT, EULER, TILE, NS>, TILE, SMEM, nfull, cap_cache, &grid);
        if (rc) return rc;
        if (grid > 0) {
            lv::so3_reparam_bwd_pipe_kernel<T, KT, EULER, TILE, NS><<<grid, TILE, SMEM, st>>>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, nfull, B, k, seed, offset);
            done = nfull * TILE;
        }
    }
    if (done < total) {
        const int64_t rest = total - done;
        const unsigned grid = unsigned((rest + TILE - 1) / TILE);
        lv::so3_reparam_bwd_kernel<T, KT, EULER, TILE><<<grid, TILE, 0, st>>>(
            mu + (done ? done * 9 : 0), sigma + (done ? done * 3 : 0), eps ? eps + done * 3 : nullptr, gz ? gz + done * 9 : nullptr,
            gangles ? gangles + done * 3 : nullptr, glq ? glq + done : nullptr, gmu + done * 9, gsigma + done * 3, rest,
            done ? rest : B, k, al, seed, offset + uint64_t(done));
    }
    return LV_OK;
}

template <typename T, bool EULER>
static int reparam_fwd(const char* name, const T* mu, const T* sigma, const T* eps, T* z, T* angles,
                       T* log_q, int64_t n, int64_t B, int k, void* stream, bool philox = false, uint64_t seed = 0, uint64_t offset = 0) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || (!eps && !philox) || (EULER ? !angles : !z)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    if (philox) eps = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int TILE = sizeof(T) == 8 ? lv::RP_TILE / 2 : lv::RP_TILE;
    if constexpr (sizeof(T) == 8) rc = reparam_fwd_launch<T, 0, EULER, TILE>(mu, sigma, eps, z, angles, log_q, n, B, k, st, seed, offset);
    else if (k == 3) rc = reparam_fwd_launch<T, 3, EULER, TILE>(mu, sigma, eps, z, angles, log_q, n, B, k, st, seed, offset);
    else if (k == 10) rc = reparam_fwd_launch<T, 10, EULER, TILE>(mu, sigma, eps, z, angles, log_q, n, B, k, st, seed, offset);
    else rc = reparam_fwd_launch<T, 0, EULER, TILE>(mu, sigma, eps, z, angles, log_q, n, B, k, st, seed, offset);
    if (rc) return rc;
    return lv::check_launch(name);
}

template <typename T, bool EULER>
static int reparam_bwd(const char* name, const T* mu, const T* sigma, const T* eps, const T* gz,
                       const T* gangles, const T* glq, T* gmu, T* gsigma, int64_t n, int64_t B, int k,
                       void* stream, bool philox = false, uint64_t seed = 0, uint64_t offset = 0) {
    int rc = reparam_check(name, n, B, k);
    if (rc) return rc;
    const int64_t total = n * B;
    if (total == 0) return LV_OK;
    if (!mu || !sigma || (!eps && !philox) || !gmu || !gsigma || (EULER && !gangles)) { lv::set_error("%s: null pointer", name); return LV_ERR_ARG; }
    if (philox) eps = nullptr;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    constexpr int TILE = sizeof(T) == 8 ? lv::RP_TILE / 2 : lv::RP_TILE;
    if constexpr (sizeof(T) == 8) rc = reparam_bwd_launch<T, 0, EULER, TILE>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, st, seed, offset);
    else if (k == 3) rc = reparam_bwd_launch<T, 3, EULER, TILE>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, st, seed, offset);
    else if (k == 10) rc = reparam_bwd_launch<T, 10, EULER, TILE>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, st, seed, offset);
    else rc = reparam_bwd_launch<T, 0, EULER, TILE>(mu, sigma, eps, gz, gangles, glq, gmu, gsigma, n, B, k, st, seed, offset);
    if (rc) return rc;
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

/* in-kernel noise: eps ~ N(0,1) is generated per sample from (seed, offset + flat sample index) -- forward and backward draw  \
   the same numbers, lv_philox_normal_* writes them out.  angles == NULL: plain reparameterize (z required);                  \
   angles != NULL: fused with matrix -> ZYZ Euler (z optional). */
#define LV_REPARAM_PHILOX_ENTRY(SFX, T)                                                                                          \
    extern "C" int lv_so3_reparam_philox_fwd_##SFX(const T* mu, const T* sigma, int64_t seed, int64_t offset, T* z, T* angles,   \
                                                   T* log_q, int64_t n, int64_t B, int k, void* stream) {                        \
        if (angles) return reparam_fwd<T, true>("so3_reparam_philox_fwd", mu, sigma, nullptr, z, angles, log_q, n, B, k, stream, \
                                                true, uint64_t(seed), uint64_t(offset));                                         \
        return reparam_fwd<T, false>("so3_reparam_philox_fwd", mu, sigma, nullptr, z, nullptr, log_q, n, B, k, stream, true,     \
                                     uint64_t(seed), uint64_t(offset));                                                          \
    }                                                                                                                            \
    extern "C" int lv_so3_reparam_philox_bwd_##SFX(const T* mu, const T* sigma, int64_t seed, int64_t offset, const T* gz,       \
                                                   const T* gangles, const T* glq, T* gmu, T* gsigma, int64_t n, int64_t B,      \
                                                   int k, void* stream) {                                                        \
        if (gangles) return reparam_bwd<T, true>("so3_reparam_philox_bwd", mu, sigma, nullptr, gz, gangles, glq, gmu, gsigma, n, \
                                                 B, k, stream, true, uint64_t(seed), uint64_t(offset));                          \
        return reparam_bwd<T, false>("so3_reparam_philox_bwd", mu, sigma, nullptr, gz, nullptr, glq, gmu, gsigma, n, B, k,       \
                                     stream, true, uint64_t(seed), uint64_t(offset));                                            \
    }                                                                                                                            \
    extern "C" int lv_philox_normal_##SFX(T* out, int64_t rows, int64_t seed, int64_t offset, void* stream) {                    \
        if (rows < 0) { lv::set_error("philox_normal: negative size"); return LV_ERR_ARG; }                                      \
        if (rows == 0) return LV_OK;                                                                                             \
        if (!out) { lv::set_error("philox_normal: null pointer"); return LV_ERR_ARG; }                                           \
        lv::philox_normal_kernel<T><<<unsigned((rows + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(           \
            out, rows, uint64_t(seed), uint64_t(offset));                                                                        \
        return lv::check_launch("philox_normal");                                                                                \
    }
LV_REPARAM_PHILOX_ENTRY(f32, float)
LV_REPARAM_PHILOX_ENTRY(f64, double)
