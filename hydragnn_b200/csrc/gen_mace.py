"""Generates hgb_mace_gen.cuh: straight-line device code for the MACE tensor-product paths and the symmetric
contraction, specialised per (lmax_in, lmax_sh) / (lmax_in, lmax_out) so that every spherical index is a compile-time
constant (accumulators stay in registers).  The numbers come from hydragnn_b200/e3.py (real Wigner 3j, U tensors).

usage: python hydragnn_b200/csrc/gen_mace.py            (build.py runs it when the header is missing or stale)
"""
import math
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

TP_CONFIGS = [(lin, lsh) for lsh in (1, 2, 3) for lin in (0, 1, 2) if lin <= lsh]
SC_CONFIGS = [(lin, lout) for lin in (1, 2, 3) for lout in (0, 1, 2) if lout <= lin]      # correlation 2 only


def fl(x):
    s = "%.9g" % x
    if "." not in s and "e" not in s:
        s += ".0"
    return s + "f"


def tp_tables(e3, lin, lsh):
    paths = e3.tp_paths(lin, lsh, lsh)
    n_p = [sum(1 for p in paths if p[2] == l) for l in range(lsh + 1)]
    acc_base, b = [], 0
    for l in range(lsh + 1):
        acc_base.append(b)
        b += n_p[l] * (2 * l + 1)
    slot, seen = [], [0] * (lsh + 1)
    for (_, _, l3) in paths:
        slot.append(seen[l3])
        seen[l3] += 1
    nnz = []
    for (l1, l2, l3) in paths:
        c = e3.w3j(l1, l2, l3) * math.sqrt(2 * l3 + 1)
        items = [(m1, m2, m3, float(c[m1, m2, m3])) for m1 in range(2 * l1 + 1) for m2 in range(2 * l2 + 1) for m3 in range(2 * l3 + 1)
                 if abs(float(c[m1, m2, m3])) > 1e-12]
        nnz.append(items)
    return paths, n_p, acc_base, slot, nnz, b


def gen_tp(e3, lin, lsh, out):
    paths, n_p, acc_base, slot, nnz, nacc = tp_tables(e3, lin, lsh)
    s_in, s_sh = (lin + 1) ** 2, (lsh + 1) ** 2
    w = out.append
    w("template <> struct MaceTP<%d, %d> {" % (lin, lsh))
    w("  static constexpr int S_IN = %d, S_SH = %d, NPATH = %d, NACC = %d, LOUT = %d;" % (s_in, s_sh, len(paths), nacc, lsh))
    w("  // per output degree l3: number of paths and first accumulator row; accumulator row = base + m3 * n_p + slot")
    w("  __host__ __device__ static constexpr int n_paths(int l3) { return %s; }" %
      " : ".join(["l3 == %d ? %d" % (l, n_p[l]) for l in range(lsh)] + [str(n_p[lsh])]))
    w("  __host__ __device__ static constexpr int acc_base(int l3) { return %s; }" %
      " : ".join(["l3 == %d ? %d" % (l, acc_base[l]) for l in range(lsh)] + [str(acc_base[lsh])]))

    def arow(k, m3):
        l3 = paths[k][2]
        return acc_base[l3] + m3 * n_p[l3] + slot[k]

    # ---- forward: acc[row][t] += (coef y[m2]) * (w[k][t] up[m1][t])
    w("  template <int CPL> __device__ __forceinline__ static void fwd(const float* y, const float (&up)[S_IN][CPL], const float (&w)[NPATH][CPL], float (&acc)[NACC][CPL]) {")
    for k, (l1, l2, l3) in enumerate(paths):
        w("    {  // path %d: %d x %d -> %d" % (k, l1, l2, l3))
        for q, (m1, m2, m3, c) in enumerate(nnz[k]):
            w("      const float c%d = %s * y[%d];" % (q, fl(c), l2 * l2 + m2))
        w("#pragma unroll")
        w("      for (int t = 0; t < CPL; ++t) {")
        for m1 in range(2 * l1 + 1):
            w("        const float u%d = w[%d][t] * up[%d][t];" % (m1, k, l1 * l1 + m1))
        for q, (m1, m2, m3, c) in enumerate(nnz[k]):
            r = arow(k, m3)
            w("        acc[%d][t] = fmaf(c%d, u%d, acc[%d][t]);" % (r, q, m1, r))
        w("      }")
        w("    }")
    w("  }")
    # ---- backward, edge side: gw[k][t] = sum c y up g ; gy[m2] += coef w up g (per-lane partial)
    w("  template <int CPL, bool NEED_Y> __device__ __forceinline__ static void bwd_edge(const float* y, const float (&up)[S_IN][CPL], const float (&w)[NPATH][CPL], const float (&g)[NACC][CPL], float (&gw)[NPATH][CPL], float (&gy)[S_SH]) {")
    for k, (l1, l2, l3) in enumerate(paths):
        w("    {  // path %d" % k)
        w("#pragma unroll")
        w("      for (int t = 0; t < CPL; ++t) {")
        w("        float a = 0.f;")
        for q, (m1, m2, m3, c) in enumerate(nnz[k]):
            r = arow(k, m3)
            w("        { const float ug = up[%d][t] * g[%d][t]; a = fmaf(%s * y[%d], ug, a); if (NEED_Y) gy[%d] = fmaf(%s * w[%d][t], ug, gy[%d]); }" %
              (l1 * l1 + m1, r, fl(c), l2 * l2 + m2, l2 * l2 + m2, fl(c), k, l2 * l2 + m2))
        w("        gw[%d][t] = a;" % k)
        w("      }")
        w("    }")
    w("  }")
    # ---- backward, sender side: gup[m1][t] += (coef y) w g
    w("  template <int CPL> __device__ __forceinline__ static void bwd_up(const float* y, const float (&w)[NPATH][CPL], const float (&g)[NACC][CPL], float (&gup)[S_IN][CPL]) {")
    for k, (l1, l2, l3) in enumerate(paths):
        w("    {  // path %d" % k)
        for q, (m1, m2, m3, c) in enumerate(nnz[k]):
            w("      const float c%d = %s * y[%d];" % (q, fl(c), l2 * l2 + m2))
        w("#pragma unroll")
        w("      for (int t = 0; t < CPL; ++t) {")
        for q, (m1, m2, m3, c) in enumerate(nnz[k]):
            r = arow(k, m3)
            w("        gup[%d][t] = fmaf(c%d * w[%d][t], g[%d][t], gup[%d][t]);" % (l1 * l1 + m1, q, k, r, l1 * l1 + m1))
        w("      }")
        w("    }")
    w("  }")
    w("};")
    w("")


def sc_tables(e3, lin, lout):
    """Correlation 2: per output degree L the nonzeros of U2[m, i, j, k] and U1[m, i, k]."""
    tabs = []
    for L in range(lout + 1):
        u2, u1 = e3.u_matrix(lin, L, 2), e3.u_matrix(lin, L, 1)
        if L == 0:
            u2, u1 = u2.unsqueeze(0), u1.unsqueeze(0)
        n2 = [(m, i, j, k, float(u2[m, i, j, k])) for m in range(u2.shape[0]) for i in range(u2.shape[1]) for j in range(u2.shape[2])
              for k in range(u2.shape[3]) if abs(float(u2[m, i, j, k])) > 1e-12]
        n1 = [(m, i, k, float(u1[m, i, k])) for m in range(u1.shape[0]) for i in range(u1.shape[1]) for k in range(u1.shape[2])
              if abs(float(u1[m, i, k])) > 1e-12]
        tabs.append((n2, n1, u2.shape[3], u1.shape[2]))
    return tabs


def gen_sc(e3, lin, lout, out):
    tabs = sc_tables(e3, lin, lout)
    s = (lin + 1) ** 2
    w = out.append
    w("template <> struct MaceSC<%d, %d> {" % (lin, lout))
    w("  static constexpr int S = %d, NOUT = %d, LOUT = %d;" % (s, (lout + 1) ** 2, lout))
    w("  // weight slots per output degree L: first P2(L) rows = weights_max (correlation 2), then P1(L) rows = weights.0")
    w("  __host__ __device__ static constexpr int p2(int L) { return %s; }" %
      " : ".join(["L == %d ? %d" % (L, tabs[L][2]) for L in range(lout)] + [str(tabs[lout][2])]))
    w("  __host__ __device__ static constexpr int p1(int L) { return %s; }" %
      " : ".join(["L == %d ? %d" % (L, tabs[L][3]) for L in range(lout)] + [str(tabs[lout][3])]))
    ktot = sum(t[2] + t[3] for t in tabs)
    w("  static constexpr int KTOT = %d;" % ktot)
    # out[m] = sum_{i} x_i ( sum_j x_j sum_k U2[m,i,j,k] w2[k] + sum_k U1[m,i,k] w1[k] )
    w("  __device__ __forceinline__ static void fwd(const float (&x)[S], const float (&wt)[KTOT], float (&o)[NOUT]) {")
    koff = 0
    for L, (n2, n1, p2, p1) in enumerate(tabs):
        for m in range(2 * L + 1):
            terms = ["%s * wt[%d] * x[%d] * x[%d]" % (fl(c), koff + k, i, j) for (mm, i, j, k, c) in n2 if mm == m]
            terms += ["%s * wt[%d] * x[%d]" % (fl(c), koff + p2 + k, i) for (mm, i, k, c) in n1 if mm == m]
            w("    o[%d] = %s;" % (L * L + m, " + ".join(terms) if terms else "0.f"))
        koff += p2 + p1
    w("  }")
    # backward: gx[i], gwt[k] from go[m]
    w("  __device__ __forceinline__ static void bwd(const float (&x)[S], const float (&wt)[KTOT], const float (&go)[NOUT], float (&gx)[S], float (&gwt)[KTOT]) {")
    w("#pragma unroll")
    w("    for (int i = 0; i < S; ++i) gx[i] = 0.f;")
    w("#pragma unroll")
    w("    for (int k = 0; k < KTOT; ++k) gwt[k] = 0.f;")
    koff = 0
    for L, (n2, n1, p2, p1) in enumerate(tabs):
        for (m, i, j, k, c) in n2:
            w("    { const float a = %s * go[%d]; gwt[%d] = fmaf(a, x[%d] * x[%d], gwt[%d]); const float b = a * wt[%d]; gx[%d] = fmaf(b, x[%d], gx[%d]); gx[%d] = fmaf(b, x[%d], gx[%d]); }" %
              (fl(c), L * L + m, koff + k, i, j, koff + k, koff + k, i, j, i, j, i, j))
        for (m, i, k, c) in n1:
            w("    { const float a = %s * go[%d]; gwt[%d] = fmaf(a, x[%d], gwt[%d]); gx[%d] = fmaf(a, wt[%d], gx[%d]); }" %
              (fl(c), L * L + m, koff + p2 + k, i, koff + p2 + k, i, koff + p2 + k, i))
        koff += p2 + p1
    w("  }")
    w("};")
    w("")


def generate(path=None):
    from hydragnn_b200 import e3
    out = ["// GENERATED by hydragnn_b200/csrc/gen_mace.py -- do not edit.  Coupling constants: hydragnn_b200/e3.py.",
           "#pragma once", "",
           "template <int LIN, int LSH> struct MaceTP;      // tensor-product paths F x (0..LIN) (x) Y(0..LSH) -> F x (0..LSH)",
           "template <int LIN, int LOUT> struct MaceSC;     // symmetric contraction, correlation 2: (0..LIN) -> (0..LOUT)", ""]
    for lin, lsh in TP_CONFIGS:
        gen_tp(e3, lin, lsh, out)
    for lin, lout in SC_CONFIGS:
        gen_sc(e3, lin, lout, out)
    text = "\n".join(out) + "\n"
    path = path or os.path.join(HERE, "hgb_mace_gen.cuh")
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


if __name__ == "__main__":
    print(generate())
