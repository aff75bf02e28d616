// Wigner-D action backward for a shared spectrum (ActionNet.item_rep, decoders.py:53) -- degree-specialised warps (sm_100a).
// Included by wigner.cu inside namespace lv (uses its trig / mbarrier / generator helpers).
//
// What bound the first TMA-fed backward (wigner_bwd_ws_kernel) was the shared-memory data pipe (81 % busy): every byte of
// the g_y tile crossed shared memory four times (TMA in, math read, spectrum-gradient write, column-sum read for the batch
// reduction) and the spectrum and trig tables were re-read per column.  Here the tile is read exactly once and never written:
//
//   * a math warp is bound to a *degree group* (e.g. {8}, {7}, {6,0}, {5,2}, {4,3,1}) and its lanes to a fixed channel
//     (lane = 10 * sample_in_slice + channel, 30 lanes = 3 samples x 10 channels; lanes 30-31 idle).  Everything that depends
//     only on (degree, channel) therefore lives in registers for the whole launch:
//       - acc_l  += g_s          the batch reduction of the item_rep gradient (lie_tools.py:251 transposed, summed over n):
//                                no tile write-back, no column-sum pass, fixed order -> bit-reproducible;
//       - U_k = G_k s, k=x,y,z   the body-frame generators applied to the spectrum (see wigner.cu), so the angle-gradient
//                                forms are three dot products T_k = <g_s, U_k> in packed FMAs instead of ~6 scalar FMAs per
//                                row, and the spectrum itself is never loaded again.
//   * an item is (tile, 3-sample slice); the warps of a group take items round-robin (static: reproducible); a tile buffer is
//     released when every (group, slice) pair has consumed it (mbarrier, count = groups x slices).
//   * per item a lane leaves only its three T_k partials in shared memory -- in three slots of the tile it has itself
//     already consumed -- and the producer warp(s) sum them over groups and channels, apply the body-frame relation and store
//     g_angles; they also own the TMA ring (bulk load of tile r + NB as soon as tile r is released) and the trig tables.
//
// Shared-memory wavefronts per sample: 81 x 10 / 30 (g_y) + 2 x 60 x 10 / 30 (trig) + ~15 vs 139 before.
#pragma once

namespace dg {

using wg2::PDeg;
using wg2::f32x2_t;

__device__ __forceinline__ f32x2_t add2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

template <int L> __device__ __forceinline__ void vzero(typename PDeg<L>::Vec& v) {
    using D = PDeg<L>;
#pragma unroll
    for (int q = 0; q < (D::NP > 0 ? D::NP : 1); ++q) { v.lo[q] = 0ull; v.hi[q] = 0ull; }
#pragma unroll
    for (int s = 0; s < (D::NS > 0 ? D::NS : 1); ++s) { v.slo[s] = 0.f; v.shi[s] = 0.f; }
    v.ctr = 0.f;
}
template <int L> __device__ __forceinline__ void vacc(typename PDeg<L>::Vec& a, const typename PDeg<L>::Vec& g) {
    using D = PDeg<L>;
#pragma unroll
    for (int q = 0; q < D::NP; ++q) { a.lo[q] = add2(a.lo[q], g.lo[q]); a.hi[q] = add2(a.hi[q], g.hi[q]); }
#pragma unroll
    for (int s = 0; s < D::NS; ++s) { a.slo[s] += g.slo[s]; a.shi[s] += g.shi[s]; }
    a.ctr += g.ctr;
}
// (ap, as) += <g, u> : packed part and scalar part
template <int L> __device__ __forceinline__ void vdot(const typename PDeg<L>::Vec& g, const typename PDeg<L>::Vec& u, f32x2_t& ap, float& as) {
    using D = PDeg<L>;
#pragma unroll
    for (int q = 0; q < D::NP; ++q) { ap = wg2::fma2(g.lo[q], u.lo[q], ap); ap = wg2::fma2(g.hi[q], u.hi[q], ap); }
#pragma unroll
    for (int s = 0; s < D::NS; ++s) { as = fmaf(g.slo[s], u.slo[s], as); as = fmaf(g.shi[s], u.shi[s], as); }
    as = fmaf(g.ctr, u.ctr, as);
}

template <int... Ls> struct DegList {};

// per-thread persistent state of one degree: batch accumulator and the three generator vectors of this lane's channel
template <int L> struct DegState { typename PDeg<L>::Vec acc, ux, uy, uz; };
template <class GL> struct GroupState;
template <> struct GroupState<DegList<>> {};
template <int L, int... R> struct GroupState<DegList<L, R...>> { DegState<L> head; GroupState<DegList<R...>> tail; };

template <class GL> struct GroupInfo;
template <> struct GroupInfo<DegList<>> { static constexpr int first = 0, elems = 0; };
template <int L, int... R> struct GroupInfo<DegList<L, R...>> {
    static constexpr int first = L;                                            // its slots carry the T_k partials (needs 2L+1 >= 3)
    static constexpr int elems = 2 * L + 1 + GroupInfo<DegList<R...>>::elems;
};

struct TAcc { f32x2_t px = 0ull, py = 0ull, pz = 0ull; float sx = 0.f, sy = 0.f, sz = 0.f; };

template <int C> __device__ __forceinline__ void group_init(GroupState<DegList<>>&, const float*) {}
template <int C, int L, int... R>
__device__ __forceinline__ void group_init(GroupState<DegList<L, R...>>& st, const float* __restrict__ spec_c) {
    using D = PDeg<L>;
    typename D::Vec s;
    vzero<L>(s);
    D::template load<true>(s, spec_c + L * L * C, C);
    vzero<L>(st.head.ux); vzero<L>(st.head.uy); vzero<L>(st.head.uz);
    D::gvecs(s, st.head.ux, st.head.uy, st.head.uz);
    vzero<L>(st.head.acc);
    group_init<C>(st.tail, spec_c);
}

// one column (sample, channel) of the tile through every degree of the group: g_s = X(c)^T J X(b)^T J X(a)^T g
template <int C> __device__ __forceinline__ void group_run(GroupState<DegList<>>&, const float*, const float4*, TAcc&) {}
template <int C, int L, int... R>
__device__ __forceinline__ void group_run(GroupState<DegList<L, R...>>& st, const float* __restrict__ col, const float4* __restrict__ tg, TAcc& t) {
    using D = PDeg<L>;
    typename D::Vec x, y;
    D::template load<false>(y, col + L * L * C, C);
    D::template xrot<true>(y, tg);
    D::jmul(y, x);
    D::template xrot<true>(x, tg + 4);
    D::jmul(x, y);
    D::template xrot<true>(y, tg + 8);
    if constexpr (L > 0) {
        vdot<L>(y, st.head.ux, t.px, t.sx);
        vdot<L>(y, st.head.uy, t.py, t.sy);
        vdot<L>(y, st.head.uz, t.pz, t.sz);
    }
    vacc<L>(st.head.acc, y);
    group_run<C>(st.tail, col, tg, t);
}

// accumulators -> red[(elem) * 32 + lane], elements in the group's degree order, natural index order inside a degree
__device__ __forceinline__ void group_flush(const GroupState<DegList<>>&, float*) {}
template <int L, int... R>
__device__ __forceinline__ void group_flush(const GroupState<DegList<L, R...>>& st, float* __restrict__ red) {
    PDeg<L>::store(st.head.acc, red, 32);
    group_flush(st.tail, red + (2 * L + 1) * 32);
}

// element m (0 .. M-1, degree-major) -> (group, index inside the group's flush order); filled on the host side of the config
template <class GL> struct GroupMap;
template <> struct GroupMap<DegList<>> { static __device__ __forceinline__ int find(int, int) { return -1; } };
template <int L, int... R> struct GroupMap<DegList<L, R...>> {
    static __device__ __forceinline__ int find(int m, int base) {
        if (m >= L * L && m < (L + 1) * (L + 1)) return base + (m - L * L);
        return GroupMap<DegList<R...>>::find(m, base + 2 * L + 1);
    }
};

// ---- configurations: degree groups, math warps per group, producer warps ------------------------------------------------
// work per column and degree (instructions: loads + 3 pair rotations + 2 J multiplies + 3 dots + accumulate):
//   l = 8: 191, 7: 179, 6: 169, 5: 124, 4: 83, 3: 75, 2: 69, 1: 36, 0: 2
struct Cfg8A {      // 15 math warps + 1 producer warp; group loads 191 / 179 / 171 / 193 / 194
    static constexpr int LT = 8, NG = 5, PROD = 1, S = 12, NB = 5, SPI = 1;
    using G0 = DegList<8>;        static constexpr int W0 = 3;
    using G1 = DegList<7>;        static constexpr int W1 = 3;
    using G2 = DegList<6, 0>;     static constexpr int W2 = 3;
    using G3 = DegList<5, 2>;     static constexpr int W3 = 3;
    using G4 = DegList<4, 3, 1>;  static constexpr int W4 = 3;
};
struct Cfg8B {      // 14 math warps + 2 producer warps; group loads per warp 64 / 60 / 68 / 66 / 76
    static constexpr int LT = 8, NG = 5, PROD = 2, S = 12, NB = 5, SPI = 1;
    using G0 = DegList<8, 0>;     static constexpr int W0 = 3;
    using G1 = DegList<7>;        static constexpr int W1 = 3;
    using G2 = DegList<6, 1>;     static constexpr int W2 = 3;
    using G3 = DegList<5, 3>;     static constexpr int W3 = 3;
    using G4 = DegList<4, 2>;     static constexpr int W4 = 2;
};
struct Cfg8C {      // 14 math + 2 producer warps, the partition a cost model fitted to Cfg8B's measured loop lengths ranks best
    static constexpr int LT = 8, NG = 5, PROD = 2, S = 12, NB = 5, SPI = 1;
    using G0 = DegList<8>;        static constexpr int W0 = 3;
    using G1 = DegList<7>;        static constexpr int W1 = 3;
    using G2 = DegList<6, 1>;     static constexpr int W2 = 3;
    using G3 = DegList<4, 3, 2>;  static constexpr int W3 = 3;
    using G4 = DegList<5, 0>;     static constexpr int W4 = 2;
};
struct Cfg8CP : Cfg8C { static constexpr int SPI = 2; };            // two slices per item
struct Cfg8BP : Cfg8B { static constexpr int SPI = 2; };
struct Cfg8BQ : Cfg8B { static constexpr int SPI = 4; };            // a whole 12-sample tile per item
struct Cfg8B18P : Cfg8B { static constexpr int S = 18, NB = 3, SPI = 2; };
struct Cfg8B18T : Cfg8B { static constexpr int S = 18, NB = 3, SPI = 3; };
// Measured at 2^20 samples (ms): Cfg8BP 0.637, Cfg8E 0.637, Cfg8F 0.717, Cfg8CP 0.661, Cfg8BQ 0.741, Cfg8B18P 0.723, Cfg8B18T 0.643.
// Timing-only ablations (-DLV_DG_ABL=1: producers skip the T sums and angle gradients, =2: math warps skip the chain):
// 0.630 and 0.494 ms -- the TMA feed alone streams at 6.9 TB/s and the producers are not the limiter; the kernel sits on the
// FMA pipe (ncu: fma pipe 65 % active, fmaheavy 62 %, math_pipe_throttle ~ long_scoreboard) at 0.82 of the HBM copy peak.
struct Cfg8E {      // 14 math + 2 producer warps, partition ranked best by the cost model with two slices per item
    static constexpr int LT = 8, NG = 5, PROD = 2, S = 12, NB = 5, SPI = 2;
    using G0 = DegList<8>;        static constexpr int W0 = 3;
    using G1 = DegList<7>;        static constexpr int W1 = 3;
    using G2 = DegList<6, 1, 0>;  static constexpr int W2 = 3;
    using G3 = DegList<5, 4>;     static constexpr int W3 = 3;
    using G4 = DegList<3, 2>;     static constexpr int W4 = 2;
};
struct Cfg8F {      // 13 math + 3 producer warps
    static constexpr int LT = 8, NG = 5, PROD = 3, S = 12, NB = 5, SPI = 2;
    using G0 = DegList<8>;        static constexpr int W0 = 3;
    using G1 = DegList<7, 1>;     static constexpr int W1 = 3;
    using G2 = DegList<6, 2>;     static constexpr int W2 = 3;
    using G3 = DegList<4, 3>;     static constexpr int W3 = 2;
    using G4 = DegList<5, 0>;     static constexpr int W4 = 2;
};
struct Cfg8A6 : Cfg8A { static constexpr int S = 6, NB = 10; };     // the same with 6-sample tiles: a finer-grained ring
struct Cfg8B6 : Cfg8B { static constexpr int S = 6, NB = 10; };
struct Cfg8B18 : Cfg8B { static constexpr int S = 18, NB = 3; };    // 18-sample tiles: a producer warp's 27 jobs fill one pass
struct Cfg6A {      // degrees 0..6 (BASELINE configs[3]): 14 math warps + 2 producer warps; loads per warp 42 / 40.5 / 39.5 / 34.5
    static constexpr int LT = 6, NG = 4, PROD = 2, S = 12, NB = 8, SPI = 1;
    using G0 = DegList<6>;        static constexpr int W0 = 4;
    using G1 = DegList<5, 1, 0>;  static constexpr int W1 = 4;
    using G2 = DegList<4, 3>;     static constexpr int W2 = 4;
    using G3 = DegList<2>;        static constexpr int W3 = 2;
    using G4 = DegList<>;         static constexpr int W4 = 0;
};

// measured (2^20 samples per launch, one B200, same box): Cfg8BP 0.638 ms (two slices per item, no lane-divergent code in the
// loop), Cfg8B18T 0.643, Cfg8CP 0.661 (the {4,3,2} group spills), Cfg8B18P 0.723, Cfg8BQ 0.741 (one item per tile: too coarse);
// with one slice per item: Cfg8C 0.685-0.688, Cfg8B 0.676-0.695, Cfg8B18 0.706, Cfg8A 0.869 (one producer warp cannot keep up:
// 36 jobs = two passes per tile), 6-sample tiles 1.22-1.28 (per-tile producer work dominates); first TMA-fed kernel 0.740
#ifndef LV_DG_CFG8
#define LV_DG_CFG8 Cfg8BP
#endif
struct Cfg6AP : Cfg6A { static constexpr int SPI = 2; };
struct Cfg6B {      // degrees 0..6 with a third producer warp: per sample the producers' work is what it is for degrees 0..8 while the
    // math is 56 % of it, so two producer warps bound the kernel (Cfg6AP 0.528 ms per 2^20 samples)
    static constexpr int LT = 6, NG = 4, PROD = 3, S = 12, NB = 8, SPI = 2;
    using G0 = DegList<6>;        static constexpr int W0 = 4;
    using G1 = DegList<5>;        static constexpr int W1 = 3;
    using G2 = DegList<4, 1, 0>;  static constexpr int W2 = 3;
    using G3 = DegList<3, 2>;     static constexpr int W3 = 3;
    using G4 = DegList<>;         static constexpr int W4 = 0;
};
struct Cfg6C : Cfg6A { static constexpr int S = 18, NB = 5, SPI = 3; };     // larger tiles: the per-tile producer chain is amortised over 18 samples
struct Cfg6D : Cfg6A { static constexpr int S = 18, NB = 5, SPI = 2; };
struct Cfg6E : Cfg6B { static constexpr int S = 24, NB = 4, SPI = 2; };
struct Cfg6F : Cfg6A { static constexpr int S = 24, NB = 4, SPI = 2; };
struct Cfg6G : Cfg6B { static constexpr int S = 18, NB = 5, SPI = 2; };
struct Cfg6H : Cfg6B { static constexpr int S = 30, NB = 3, SPI = 2; };
// degrees 0..6, ms per 2^20 samples: Cfg6AP 0.527, Cfg6B 0.524 (a third producer warp does not help: the per-TILE chain of the
// producers -- wait, T sums, fence, TMA issue -- is what bounds it; with the math removed 0.423), 18-sample tiles Cfg6D 0.412 /
// Cfg6C 0.421 / Cfg6G 0.420, 24-sample Cfg6E 0.408 / Cfg6F 0.461, 30-sample Cfg6H 0.405.  Cfg6D = 0.77 of the HBM copy peak.
#ifndef LV_DG_CFG6
#define LV_DG_CFG6 Cfg6D
#endif

constexpr int DG_C = 10, DG_SL = 3;          // channels; samples per slice (30 lanes)

template <class CFG> struct Geo {
    static constexpr int M = (CFG::LT + 1) * (CFG::LT + 1), MC = M * DG_C;
    static constexpr int SLICES = CFG::S / DG_SL;
    static constexpr int SPI = CFG::SPI;                        // slices per work item
    static constexpr int MATH = CFG::W0 + CFG::W1 + CFG::W2 + CFG::W3 + CFG::W4;
    static constexpr int WARPS = MATH + CFG::PROD, THREADS = WARPS * 32;
    static constexpr uint32_t TILE_BYTES = CFG::S * MC * 4u;
    static constexpr int TILE_FLOATS = CFG::S * MC;
    static constexpr int SPW = CFG::S / CFG::PROD;              // samples per producer warp
    static constexpr int JOBS = SPW * 3;                        // (sample, angle) jobs per producer warp and tile
    static constexpr int PASSES = (JOBS + 29) / 30;             // 30 jobs per pass: a sample's three jobs stay in one warp pass
    static constexpr int RED_STRIDE = 24;                       // elements per warp in the final reduction buffer (>= max group elems)
    static constexpr size_t SMEM = size_t(CFG::NB) * TILE_FLOATS * 4 + size_t(CFG::NB) * CFG::S * WG_TRIG_STRIDE * 4 + 2 * CFG::NB * 8 +
                                   size_t(CFG::PROD) * PASSES * 32 * 4 + 16;
    static_assert(CFG::S % DG_SL == 0 && CFG::S % CFG::PROD == 0 && (CFG::S * MC) % 4 == 0 && SLICES % SPI == 0, "tile geometry");
    static_assert(WARPS == 16, "16 warps x 128 registers fill the register file");
    static_assert(GroupInfo<typename CFG::G0>::elems <= RED_STRIDE && GroupInfo<typename CFG::G1>::elems <= RED_STRIDE &&
                  GroupInfo<typename CFG::G2>::elems <= RED_STRIDE && GroupInfo<typename CFG::G3>::elems <= RED_STRIDE &&
                  GroupInfo<typename CFG::G4>::elems <= RED_STRIDE, "RED_STRIDE");
    static_assert(GroupInfo<typename CFG::G0>::elems + GroupInfo<typename CFG::G1>::elems + GroupInfo<typename CFG::G2>::elems +
                  GroupInfo<typename CFG::G3>::elems + GroupInfo<typename CFG::G4>::elems == M, "the groups partition the degrees");
    static_assert(size_t(MATH) * RED_STRIDE * 32 * 4 <= size_t(CFG::NB) * TILE_FLOATS * 4, "reduction buffer fits the tile ring");
};

// ---------------------------------------------------------------------------------------------------------- math warps
template <class CFG, class GL, int W>
__device__ __forceinline__ void math_group(int wg, int red_warp, const float* __restrict__ spectrum, float* tiles, const float* trig_all,
                                           uint64_t* full, uint64_t* empty, int my_tiles, float* red) {
    using G = Geo<CFG>;
    constexpr int C = DG_C, MC = G::MC, SLOT = GroupInfo<GL>::first * GroupInfo<GL>::first;
    static_assert(GroupInfo<GL>::first >= 1, "the group's first degree provides the three T_k slots");
    const int lane = threadIdx.x & 31;
    const int la = lane < 30 ? lane : 29;         // lanes 30, 31 shadow lane 29 (no divergence), their results are dropped
    const int sl = la / C, c = la - sl * C;
    GroupState<GL> st;
    group_init<C>(st, spectrum + c);
    // an item = SPI consecutive 3-sample slices of one tile: one full-barrier wait, one release and one round of bookkeeping per
    // item (SPI = 2: ~20 instructions less per slice and half the barrier traffic)
    constexpr int SPI = G::SPI, IPT = G::SLICES / SPI;
    int q = 0, p = wg, buf = 0;
    uint32_t par = 0;
    while (p >= IPT) { p -= IPT; ++q; if (++buf == CFG::NB) { buf = 0; par ^= 1u; } }
    // No lane-divergent code in the loop: lanes 30 / 31 recompute lane 29's column (same addresses, same values; their
    // accumulators are never read), and the rows a ragged last tile does not have are zero-filled by the producer (zero g_y
    // contributes zero to every sum), so no sample-validity test is needed here.
    while (q < my_tiles) {
        mbar_wait(full + buf, par);
#pragma unroll 1          // unrolling the item loop doubles the live state and spills: 1.24 ms
        for (int h = 0; h < SPI; ++h) {
            const int s = (p * SPI + h) * DG_SL + sl;
            float* col = tiles + buf * G::TILE_FLOATS + s * MC + c;
            TAcc t;
#if defined(LV_DG_ABL) && (LV_DG_ABL & 2)
            t.sx = col[0];
#else
            group_run<C>(st, col, reinterpret_cast<const float4*>(trig_all + (buf * CFG::S + s) * WG_TRIG_STRIDE), t);
#endif
            // T_k partials of this (group, column) into three slots of the tile this lane has consumed itself
            col[(SLOT + 0) * C] = t.sx + (wg2::plo(t.px) + wg2::phi(t.px));
            col[(SLOT + 1) * C] = t.sy + (wg2::plo(t.py) + wg2::phi(t.py));
            col[(SLOT + 2) * C] = t.sz + (wg2::plo(t.pz) + wg2::phi(t.pz));
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + buf);       // release: slot writes visible to the producer
        p += W;
        while (p >= IPT) { p -= IPT; ++q; if (++buf == CFG::NB) { buf = 0; par ^= 1u; } }
    }
    named_bar_sync(3, G::THREADS);                     // every tile consumed, every T slot read: the ring is free
    group_flush(st, red + red_warp * G::RED_STRIDE * 32 + lane);
}

// ---------------------------------------------------------------------------------------------------------- kernel
template <class CFG>
__global__ void __launch_bounds__(Geo<CFG>::THREADS, 1)
wigner_bwd_dg_kernel(const float* __restrict__ angles, const float* __restrict__ spectrum, const float* __restrict__ gout,
                     float* __restrict__ gangles, float* __restrict__ partial, int64_t ntiles, int last_rows, int transpose) {
    using G = Geo<CFG>;
    constexpr int C = DG_C, MC = G::MC, S = CFG::S, NB = CFG::NB, PROD = CFG::PROD;
    extern __shared__ __align__(16) float smem[];
    float* tiles = smem;                                              // [NB][S][MC]
    float* trig_all = tiles + NB * G::TILE_FLOATS;                    // [NB][S][52]
    uint64_t* full = reinterpret_cast<uint64_t*>(trig_all + NB * S * WG_TRIG_STRIDE);      // [NB] count 1 + PROD
    uint64_t* empty = full + NB;                                                           // [NB] count NG * SLICES
    float* ang_stage = reinterpret_cast<float*>(empty + NB);                               // [PROD][PASSES][32] angles of the next tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < ntiles ? int((ntiles - first + stride - 1) / stride) : 0;
    // CTA-local index of the (possibly ragged) last tile of the launch, -1 if another CTA owns it
    const int ragged_q = (my_tiles > 0 && first + int64_t(my_tiles - 1) * stride == ntiles - 1 && last_rows < S) ? my_tiles - 1 : -1;
    if (tid == 0) {
        for (int b = 0; b < NB; ++b) { mbar_init(full + b, 1 + PROD); mbar_init(empty + b, CFG::NG * (G::SLICES / G::SPI)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float* red = tiles;                                               // final reduction buffer [MATH][RED_STRIDE][32] (reuses the ring)

    constexpr int B1 = CFG::W0, B2 = B1 + CFG::W1, B3 = B2 + CFG::W2, B4 = B3 + CFG::W3, B5 = B4 + CFG::W4;
#define DG_ARGS spectrum, tiles, trig_all, full, empty, my_tiles, red
    if (warp < B1) {
        math_group<CFG, typename CFG::G0, CFG::W0>(warp, warp, DG_ARGS);
    } else if (warp < B2) {
        math_group<CFG, typename CFG::G1, CFG::W1>(warp - B1, warp, DG_ARGS);
    } else if (warp < B3) {
        math_group<CFG, typename CFG::G2, CFG::W2>(warp - B2, warp, DG_ARGS);
    } else if (warp < B4) {
        math_group<CFG, typename CFG::G3, CFG::W3>(warp - B3, warp, DG_ARGS);
    } else if (warp < B5) {
        if constexpr (CFG::W4 > 0) math_group<CFG, typename CFG::G4, CFG::W4>(warp - B4, warp, DG_ARGS);
    } else {
#undef DG_ARGS
        // ------------------------------------------------------------ producer warp(s)
        const int pw = warp - B5;
        constexpr int PASSES = G::PASSES;
        // job (pass k, lane): (sample, angle) = (pw * SPW + j / 3, j % 3), j = 30 k + lane < JOBS
        int job_s[PASSES], job_a[PASSES];
        bool job_on[PASSES];
#pragma unroll
        for (int k = 0; k < PASSES; ++k) {
            const int j = 30 * k + lane;
            job_on[k] = lane < 30 && j < G::JOBS;
            const int jj = job_on[k] ? j : 0;
            job_s[k] = pw * G::SPW + jj / 3;
            job_a[k] = jj - 3 * (jj / 3);
        }
        auto tile_rows = [&](int j) -> int { return j == ragged_q ? last_rows : S; };
        // angles travel global -> shared by cp.async (one 4-byte piece per job) a whole iteration before they are needed:
        // an LDG would be an outstanding load at the proxy fence below, which waits for it (ncu: the DRAM latency of the
        // angle fetch sat between a buffer becoming free and its refill)
        auto job_valid = [&](int j, int k) -> bool { return job_on[k] && j < my_tiles && job_s[k] < tile_rows(j); };
        auto fetch_phi = [&](int j) {
#pragma unroll
            for (int k = 0; k < PASSES; ++k)
                if (job_valid(j, k)) {
                    const int64_t n = (first + int64_t(j) * stride) * S + job_s[k];
                    cp_async4(ang_stage + (pw * PASSES + k) * 32 + lane, angles + n * 3 + (transpose ? 2 - job_a[k] : job_a[k]));
                }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto read_phi = [&](int j, float* phi) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                const float v = ang_stage[(pw * PASSES + k) * 32 + lane];
                phi[k] = job_valid(j, k) ? (transpose ? -v : v) : 0.f;
            }
        };
        auto trig_store = [&](int buf, int k, const float* tr) {
            if (job_on[k]) {
                float4* d = reinterpret_cast<float4*>(trig_all + (buf * S + job_s[k]) * WG_TRIG_STRIDE + job_a[k] * WG_TRIG_ANGLE);
#pragma unroll
                for (int i = 0; i < 4; ++i) d[i] = make_float4(tr[4 * i], tr[4 * i + 1], tr[4 * i + 2], tr[4 * i + 3]);
            }
        };
        auto zero_tail = [&](int j) {       // every producer lane: rows a ragged tile does not have become zeros (before the arrive on full)
            if (j == ragged_q) {
                float* tail = tiles + (j % NB) * G::TILE_FLOATS + last_rows * MC;
                for (int i = pw * 32 + lane; i < (S - last_rows) * MC; i += PROD * 32) tail[i] = 0.f;
            }
        };
        auto issue_load = [&](int j) {      // one lane: first arrival on full + the bulk load; L2 prefetch of the tile after it
            const int buf = j % NB;
            const uint32_t bytes = uint32_t(tile_rows(j)) * uint32_t(MC) * 4u;
            mbar_expect_tx(full + buf, bytes);
            tma_load(tiles + buf * G::TILE_FLOATS, gout + (first + int64_t(j) * stride) * S * MC, bytes, full + buf);
            if (j + 1 < my_tiles)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                             :: "l"(gout + (first + int64_t(j + 1) * stride) * S * MC), "r"(uint32_t(tile_rows(j + 1)) * uint32_t(MC) * 4u) : "memory");
        };
        float phi[PASSES];
        auto trig_tile = [&](int buf) {                 // trig tables of this lane's jobs from phi -> the buffer's table; arrive on full
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                float tr[WG_TRIG_ANGLE];
                trig_fill(tr, phi[k]);
                trig_store(buf, k, tr);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(full + buf);
        };
        // prologue: fill the ring
        fetch_phi(0);
        for (int j = 0; j < NB && j < my_tiles; ++j) {
            if (pw == 0 && lane == 0) issue_load(j);
            zero_tail(j);
            read_phi(j, phi);
            fetch_phi(j + 1);
            trig_tile(j);
        }
        // the angles of tile min(NB, my_tiles) are on their way.  Per released tile r: (1) read what the angle gradients need from the
        // buffer, (2) refill it -- this is all that sits between a buffer becoming free and its bulk load --, then off that
        // critical path (3) the trig table of the incoming tile (its angles were fetched a whole iteration ago), (4) the angle
        // gradients of tile r, (5) fetch the angles of the tile after the incoming one.
        int buf = 0;
        uint32_t par = 0;
        for (int r = 0; r < my_tiles; ++r) {
            const int jn = r + NB;
            const int rows = tile_rows(r);
            mbar_wait(empty + buf, par);
            // T_k of (sample, k = job_a) summed over the groups and channels (fixed order), and the sample's cos / sin of the
            // second and third effective angles
            const float* tile = tiles + buf * G::TILE_FLOATS;
            float tk[PASSES], cb[PASSES], sb[PASSES], cc[PASSES], sc[PASSES];
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                float a0[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, a1[5] = {0.f, 0.f, 0.f, 0.f, 0.f};     // one chain per group: short dependency chains
#if defined(LV_DG_ABL) && (LV_DG_ABL & 1)
                if (false) {
#else
                if (job_on[k] && job_s[k] < rows) {
#endif
                    const float* base = tile + job_s[k] * MC + job_a[k] * C;
                    constexpr int slots[5] = {GroupInfo<typename CFG::G0>::first, GroupInfo<typename CFG::G1>::first, GroupInfo<typename CFG::G2>::first,
                                              GroupInfo<typename CFG::G3>::first, CFG::W4 > 0 ? GroupInfo<typename CFG::G4>::first : 0};
#pragma unroll
                    for (int g = 0; g < CFG::NG; ++g) {
                        const float2* p2 = reinterpret_cast<const float2*>(base + slots[g] * slots[g] * C);
#pragma unroll
                        for (int i = 0; i < C / 2; ++i) { const float2 v = p2[i]; a0[g] += v.x; a1[g] += v.y; }
                    }
                    const float* tr_s = trig_all + (buf * S + job_s[k]) * WG_TRIG_STRIDE;
                    cb[k] = tr_s[WG_TRIG_ANGLE]; sb[k] = tr_s[WG_TRIG_ANGLE + 2];
                    cc[k] = tr_s[2 * WG_TRIG_ANGLE]; sc[k] = tr_s[2 * WG_TRIG_ANGLE + 2];
                } else {
                    cb[k] = sb[k] = cc[k] = sc[k] = 0.f;
                }
                tk[k] = ((a0[0] + a1[0]) + (a0[1] + a1[1])) + ((a0[2] + a1[2]) + (a0[3] + a1[3])) + (a0[4] + a1[4]);
            }
            if (jn < my_tiles) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy accesses of the buffer before the bulk write
                if constexpr (PROD > 1) named_bar_sync(1, PROD * 32); else __syncwarp();
                if (pw == 0 && lane == 0) issue_load(jn);
                zero_tail(jn);
                read_phi(jn, phi);
                fetch_phi(jn + 1);
                trig_tile(buf);
            }
            // body-frame relation (wigner.cu): the three T_k of a sample sit in three adjacent lanes of the same pass
#pragma unroll
            for (int k = 0; k < PASSES; ++k) {
                const int b0 = lane - job_a[k];
                const float tx = __shfl_sync(0xffffffffu, tk[k], b0), ty = __shfl_sync(0xffffffffu, tk[k], b0 + 1),
                            tz = __shfl_sync(0xffffffffu, tk[k], b0 + 2);
                if (job_on[k] && job_s[k] < rows) {
                    const float gval = angle_grad_from_generators(transpose ? 2 - job_a[k] : job_a[k], tx, ty, tz, cb[k], sb[k], cc[k], sc[k]);
                    const int64_t n = (first + int64_t(r) * stride) * S + job_s[k];
                    gangles[n * 3 + job_a[k]] = transpose ? -gval : gval;
                }
            }
            if (++buf == NB) { buf = 0; par ^= 1u; }
        }
        named_bar_sync(3, G::THREADS);                 // pairs with the math warps' barrier before the ring is reused
    }
    __syncthreads();                                   // accumulators are in red
    // partial row of this CTA: element (m, c) summed over the group's warps and the three sample lanes, fixed order
    for (int o = tid; o < MC; o += G::THREADS) {
        const int m = o / C, c = o - m * C;
        int w0, nw, e;
        if ((e = GroupMap<typename CFG::G0>::find(m, 0)) >= 0) { w0 = 0; nw = CFG::W0; }
        else if ((e = GroupMap<typename CFG::G1>::find(m, 0)) >= 0) { w0 = B1; nw = CFG::W1; }
        else if ((e = GroupMap<typename CFG::G2>::find(m, 0)) >= 0) { w0 = B2; nw = CFG::W2; }
        else if ((e = GroupMap<typename CFG::G3>::find(m, 0)) >= 0) { w0 = B3; nw = CFG::W3; }
        else { e = GroupMap<typename CFG::G4>::find(m, 0); w0 = B4; nw = CFG::W4; }
        float a = 0.f;
        for (int w = 0; w < nw; ++w)
#pragma unroll
            for (int s = 0; s < DG_SL; ++s) a += red[((w0 + w) * G::RED_STRIDE + e) * 32 + s * C + c];
        partial[int64_t(blockIdx.x) * MC + o] = a;
    }
}

}  // namespace dg
