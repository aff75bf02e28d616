#!/usr/bin/env python
"""Generate lie_vae_b200/csrc/wigner_gen.cuh: per-degree Wigner chain operators in packed f32x2 form.

Blackwell (sm_100a) has two-wide FP32 instructions (PTX fma/mul.rn.f32x2, SASS FFMA2 / FMUL2) that do two
independent FMAs per issue slot on an aligned register pair.  The FP32 pipe does not get faster, but the
Wigner kernels are *issue* bound, so halving the number of FP instructions is what counts.  Instead of
giving a thread two channels (twice the registers, measured slower), the 2l+1 entries of ONE column are
packed pairwise so that every operator of the chain  X(a) J X(b) J X(c)  works on whole pairs:

  * entries are addressed by frequency m: lo_m = x[l-m], hi_m = x[l+m], centre x[l];
  * frequencies of equal parity are paired: (1,3) (5,7) | (2,4) (6,8); a class with an odd count leaves one
    frequency unpaired ("single"), which stays scalar;
  * X(phi) rotates (lo_m, hi_m) by m*phi: on a pair of frequencies that is 2 FMUL2 + 2 FFMA2 with the
    (cos, cos) and (sin, sin) pairs read from the trig table (negations fold into operand modifiers);
  * J_l (Pinchon-Hoggan, ~25 % dense) is block structured by exactly these classes (rows lo/even, lo/odd,
    hi/odd, centre+hi/even each read one class), so two paired output rows share their input columns:
    (y_r1, y_r2) += (J[r1,c], J[r2,c]) * broadcast(x_c) is ONE FFMA2 whose coefficient pair comes from
    constant memory through a uniform register pair and whose x operand is a scalar-broadcast register;
  * the generator products <h, G w> pair the same way.

Usage:  python tools/gen_wigner.py [LMAX]      (default 8)
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("jmatrix", os.path.join(ROOT, "lie_vae_b200", "jmatrix.py"))
jm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(jm)


def lit(v):
    s = "%.9g" % v
    return s + "f" if ("." in s or "e" in s) else s + ".0f"


def layout(l):
    """pair slots [(m1, m2)], single frequencies [m], and their trig-table positions."""
    odd = list(range(1, l + 1, 2))
    even = list(range(2, l + 1, 2))
    pairs, singles = [], []
    for cls in (odd, even):
        for i in range(0, len(cls) - 1, 2):
            pairs.append((cls[i], cls[i + 1]))
        if len(cls) % 2:
            singles.append(cls[-1])
    return pairs, singles


def trig_pos(m):
    """(float4 slot, lane) of frequency m in the per-angle trig table [(c_m1, c_m2, s_m1, s_m2)] x 4."""
    if m % 2:
        return (m - 1) // 4, ((m - 1) // 2) % 2
    return 2 + (m - 2) // 4, ((m - 2) // 2) % 2


class Deg:
    def __init__(self, l):
        self.l = l
        self.pairs, self.singles = layout(l)
        self.where = {}          # element index -> accessor expression template with '{v}'
        for q, (m1, m2) in enumerate(self.pairs):
            self.where[l - m1] = "plo({v}.lo[%d])" % q
            self.where[l - m2] = "phi({v}.lo[%d])" % q
            self.where[l + m1] = "plo({v}.hi[%d])" % q
            self.where[l + m2] = "phi({v}.hi[%d])" % q
        for s, m in enumerate(self.singles):
            self.where[l - m] = "{v}.slo[%d]" % s
            self.where[l + m] = "{v}.shi[%d]" % s
        self.where[l] = "{v}.ctr"

    def acc(self, i, v):
        return self.where[i].format(v=v)


def emit(l, ktab):
    d = Deg(l)
    J = jm.j_matrix_np(l)
    n = 2 * l + 1
    NP, NS = len(d.pairs), len(d.singles)
    o = []
    o.append("template <> struct PDeg<%d> {" % l)
    o.append("    static constexpr int NP = %d, NS = %d;" % (NP, NS))
    o.append("    struct Vec { f32x2_t lo[%d], hi[%d]; float slo[%d], shi[%d]; float ctr; };" % (max(NP, 1), max(NP, 1), max(NS, 1), max(NS, 1)))
    # ---- load / store (column with element stride st)
    o.append("    template <bool G> static __device__ __forceinline__ void load(Vec& v, const float* __restrict__ s, int st) {")
    for q, (m1, m2) in enumerate(d.pairs):
        o.append("        v.lo[%d] = pk(ldf<G>(s + %d * st), ldf<G>(s + %d * st));" % (q, l - m1, l - m2))
        o.append("        v.hi[%d] = pk(ldf<G>(s + %d * st), ldf<G>(s + %d * st));" % (q, l + m1, l + m2))
    for s, m in enumerate(d.singles):
        o.append("        v.slo[%d] = ldf<G>(s + %d * st); v.shi[%d] = ldf<G>(s + %d * st);" % (s, l - m, s, l + m))
    o.append("        v.ctr = ldf<G>(s + %d * st);" % l)
    o.append("    }")
    o.append("    static __device__ __forceinline__ void store(const Vec& v, float* __restrict__ s, int st) {")
    for i in range(n):
        o.append("        s[%d * st] = %s;" % (i, d.acc(i, "v")))
    o.append("    }")
    # ---- pair rotations.  t = trig table of one angle: 4 x float4 (c_m1, c_m2, s_m1, s_m2)
    o.append("    template <bool TR> static __device__ __forceinline__ void xrot(Vec& v, const float4* __restrict__ t) {")
    slots = sorted(set([trig_pos(m1)[0] for m1, _ in d.pairs] + [trig_pos(m)[0] for m in d.singles]))
    for sl in slots:
        o.append("        const float4 t%d = t[%d];" % (sl, sl))
    for q, (m1, m2) in enumerate(d.pairs):
        sl, lane = trig_pos(m1)
        assert lane == 0 and trig_pos(m2) == (sl, 1)
        o.append("        { const f32x2_t C = pk(t%d.x, t%d.y), S = pk(t%d.z, t%d.w), P = v.lo[%d], Q = v.hi[%d];" % (sl, sl, sl, sl, q, q))
        o.append("          const f32x2_t sq = mul2(S, Q), sp = mul2(S, P);")
        o.append("          v.lo[%d] = fma2(C, P, TR ? neg2(sq) : sq); v.hi[%d] = fma2(C, Q, TR ? sp : neg2(sp)); }" % (q, q))
    for s, m in enumerate(d.singles):
        sl, lane = trig_pos(m)
        c = "t%d.%s" % (sl, "xy"[lane])
        sn = "t%d.%s" % (sl, "zw"[lane])
        o.append("        { const float c = %s, s = TR ? -%s : %s, a = v.slo[%d], b = v.shi[%d];" % (c, sn, sn, s, s))
        o.append("          v.slo[%d] = fmaf(c, a, s * b); v.shi[%d] = fmaf(c, b, -(s * a)); }" % (s, s))
    if not slots:
        o.append("        (void)v; (void)t;")
    o.append("    }")
    # ---- y = J x
    o.append("    static __device__ __forceinline__ void jmul(const Vec& x, Vec& y) {")
    n_f2 = n_f1 = 0

    def scalar_row(r):
        nonlocal n_f1
        terms = [(c, J[r, c]) for c in range(n) if J[r, c] != 0.0]
        assert terms
        terms.sort(key=lambda t: 0 if abs(abs(t[1]) - 1.0) < 1e-15 else 1)
        c0, v0 = terms[0]
        if abs(v0 - 1.0) < 1e-15:
            e = d.acc(c0, "x")
        elif abs(v0 + 1.0) < 1e-15:
            e = "-" + d.acc(c0, "x")
        else:
            e = "%s * %s" % (lit(v0), d.acc(c0, "x"))
        n_f1 += len(terms)
        for c, v in terms[1:]:
            e = "fmaf(%s, %s, %s)" % (lit(v), d.acc(c, "x"), e)
        return e

    def pair_rows(r1, r2):
        nonlocal n_f2
        cols = [c for c in range(n) if J[r1, c] != 0.0 or J[r2, c] != 0.0]
        s1 = sum(1 for c in cols if J[r1, c] != 0.0)
        s2 = sum(1 for c in cols if J[r2, c] != 0.0)
        assert len(cols) <= max(s1, s2) + 1, "rows %d,%d of J_%d do not share their support" % (r1, r2, l)
        e = None
        for c in cols:
            k = len(ktab)
            ktab.append((J[r1, c], J[r2, c]))
            xb = "bc(%s)" % d.acc(c, "x")
            e = "mul2(kj(%d), %s)" % (k, xb) if e is None else "fma2(kj(%d), %s, %s)" % (k, xb, e)
            n_f2 += 1
        return e

    for q, (m1, m2) in enumerate(d.pairs):
        o.append("        y.lo[%d] = %s;" % (q, pair_rows(l - m1, l - m2)))
        o.append("        y.hi[%d] = %s;" % (q, pair_rows(l + m1, l + m2)))
    for s, m in enumerate(d.singles):
        o.append("        y.slo[%d] = %s;" % (s, scalar_row(l - m)))
        o.append("        y.shi[%d] = %s;" % (s, scalar_row(l + m)))
    o.append("        y.ctr = %s;" % scalar_row(l))
    o.append("    }")
    # ---- <h, G w> = sum_m m (h[l-m] w[l+m] - h[l+m] w[l-m]), accumulated into (ap: per-lane pair, as: scalar)
    o.append("    static __device__ __forceinline__ void gdot(const Vec& h, const Vec& w, f32x2_t& ap, float& as) {")
    for q, (m1, m2) in enumerate(d.pairs):
        o.append("        ap = fma2(pk(%s, %s), fma2(neg2(h.hi[%d]), w.lo[%d], mul2(h.lo[%d], w.hi[%d])), ap);" % (lit(m1), lit(m2), q, q, q, q))
    for s, m in enumerate(d.singles):
        o.append("        as = fmaf(%s, fmaf(h.slo[%d], w.shi[%d], -(h.shi[%d] * w.slo[%d])), as);" % (lit(m), s, s, s, s))
    if not d.pairs and not d.singles:
        o.append("        (void)h; (void)w; (void)ap; (void)as;")
    elif not d.pairs:
        o.append("        (void)ap;")
    elif not d.singles:
        o.append("        (void)as;")
    o.append("    }")
    # ---- T_x += <g, G_x s>, T_y += <g, G_y s> with the body-frame generators G_y = J G_z J, G_x = [G_z, G_y] of the chain
    #      (sparse: each row couples a frequency with its neighbours; coefficients as immediates)
    import numpy as np
    Gz = np.zeros((n, n))
    for i in range(n):
        if i != l:
            Gz[i, 2 * l - i] = l - i
    Gy = J @ Gz @ J
    Gx = Gz @ Gy - Gy @ Gz
    o.append("    static __device__ __forceinline__ void gxy_dots(const Vec& g, const Vec& s, float& tx, float& ty) {")
    n_gen = 0
    for name, G in (("tx", Gx), ("ty", Gy)):
        for i in range(n):
            terms = [(j, G[i, j]) for j in range(n) if abs(G[i, j]) > 1e-12]
            if not terms:
                continue
            e = "%s * %s" % (lit(terms[0][1]), d.acc(terms[0][0], "s"))
            for j, v in terms[1:]:
                e = "fmaf(%s, %s, %s)" % (lit(v), d.acc(j, "s"), e)
            o.append("        %s = fmaf(%s, %s, %s);" % (name, d.acc(i, "g"), e, name))
            n_gen += len(terms) + 1
    if n_gen == 0:
        o.append("        (void)g; (void)s; (void)tx; (void)ty;")
    o.append("    }")
    # ---- U_k = G_k s, k = x, y, z, as packed vectors: for a spectrum shared by every sample they are computed once per
    #      thread, and the generator forms become plain dot products T_k = <g_s, U_k> (wigner_bwd_dg.cuh)
    o.append("    static __device__ __forceinline__ void gvecs(const Vec& s, Vec& ux, Vec& uy, Vec& uz) {")
    for name, G in (("x", Gx), ("y", Gy), ("z", Gz)):
        for i in range(n):
            terms = [(j, G[i, j]) for j in range(n) if abs(G[i, j]) > 1e-12]
            if not terms:
                e = "0.0f"
            else:
                e = "%s * %s" % (lit(terms[0][1]), d.acc(terms[0][0], "s"))
                for j, v in terms[1:]:
                    e = "fmaf(%s, %s, %s)" % (lit(v), d.acc(j, "s"), e)
            o.append("        const float %s%d = %s;" % (name, i, e))
        for q, (m1, m2) in enumerate(d.pairs):
            o.append("        u%s.lo[%d] = pk(%s%d, %s%d); u%s.hi[%d] = pk(%s%d, %s%d);"
                     % (name, q, name, l - m1, name, l - m2, name, q, name, l + m1, name, l + m2))
        for si, m in enumerate(d.singles):
            o.append("        u%s.slo[%d] = %s%d; u%s.shi[%d] = %s%d;" % (name, si, name, l - m, name, si, name, l + m))
        o.append("        u%s.ctr = %s%d;" % (name, name, l))
    if n == 1:
        o.append("        (void)s;")
    o.append("    }")
    o.append("};")
    return "\n".join(o), n_f2, n_f1


HEADER = '''// GENERATED by tools/gen_wigner.py %(lmax)d -- do not edit by hand.
// Per-degree operators of the Wigner chain  X(a) J X(b) J X(c)  on ONE column, in packed f32x2 form
// (SASS FFMA2 / FMUL2: two FMAs per issue slot).  See the generator's docstring for the layout:
// lo_m = x[l-m], hi_m = x[l+m]; frequencies of equal parity are paired (1,3) (5,7) | (2,4) (6,8);
// unpaired frequencies and the centre stay scalar.  J_l coefficients of paired rows sit in constant memory
// (kJ2, read through uniform registers); coefficients of scalar rows are immediates.
#pragma once
namespace lv { namespace wg2 {
constexpr int kGenLmax = %(lmax)d;
typedef unsigned long long f32x2_t;   // two floats in an aligned 64-bit register pair
__device__ __forceinline__ f32x2_t pk(float a, float b) { f32x2_t r; asm("mov.b64 %%0, {%%1, %%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float plo(f32x2_t v) { float a, b; asm("mov.b64 {%%0, %%1}, %%2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float phi(f32x2_t v) { float a, b; asm("mov.b64 {%%0, %%1}, %%2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f32x2_t bc(float a) { return pk(a, a); }                       // -> scalar-broadcast operand (R.F32)
__device__ __forceinline__ f32x2_t neg2(f32x2_t v) { return pk(-plo(v), -phi(v)); }      // folds into an operand modifier
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %%0, %%1, %%2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %%0, %%1, %%2, %%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
template <bool G> __device__ __forceinline__ float ldf(const float* p) { return G ? __ldg(p) : *p; }
'''


def main():
    lmax = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ktab = []
    bodies, stats = [], []
    for l in range(lmax + 1):
        code, f2, f1 = emit(l, ktab)
        bodies.append(code)
        stats.append((f2, f1))
    out = [HEADER % {"lmax": lmax}]
    out.append("// (J[r1,c], J[r2,c]) for the paired rows, in order of use")
    out.append("__constant__ __align__(16) float2 kJ2[%d] = {" % max(len(ktab), 1))
    for i in range(0, len(ktab), 4):
        out.append("    " + " ".join("{%s, %s}," % (lit(a), lit(b)) for a, b in ktab[i:i + 4]))
    out.append("};")
    out.append("__device__ __forceinline__ f32x2_t kj(int i) { return pk(kJ2[i].x, kJ2[i].y); }")
    out.append("template <int L> struct PDeg;")
    out.extend(bodies)
    out.append("// J multiply, (packed FMAs, scalar FMAs) per degree: %s; issue slots %d (scalar form: 247)"
               % (stats, sum(a + b for a, b in stats)))
    out.append("}}  // namespace lv::wg2")
    path = os.path.join(ROOT, "lie_vae_b200", "csrc", "wigner_gen.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    print("wrote", path, stats, "constants", len(ktab))


if __name__ == "__main__":
    main()
