#!/usr/bin/env python
"""Straight-line forward-DFT codelet generator for the sm_100a FFT kernels.

Emits `deep_cine_cardiac_mri_b200/csrc/codelets.cuh`: one fully unrolled
`dftN(float (&re)[N], float (&im)[N])` per size, natural order in and out,
forward sign (e^{-2 pi i jk/N}), every twiddle a literal constant so ptxas
folds it into FMUL/FFMA immediates.  The inverse transform is obtained by the
callers swapping the re/im arrays (IDFT(x) = swap(DFT(swap(x)))), so only
forward codelets exist.

The generator builds a hash-consed DAG of real operations (add, sub, mul by
constant, fma with constant), tries every ordering of the mixed-radix
Cooley-Tukey factorisation over primitive butterflies {2,3,4,5}, keeps the
cheapest, dead-code-eliminates and prints it.

    python tools/gen_codelets.py            # rewrites codelets.cuh
"""
from __future__ import annotations

import itertools
import math
from pathlib import Path

OUT = Path(__file__).resolve().parents[1] / "deep_cine_cardiac_mri_b200" / "csrc" / "codelets.cuh"
SIZES = [2, 3, 4, 5, 8, 10, 12, 15, 16, 20, 24, 25, 30, 32, 40]
EPS = 1e-12


class Dag:
    def __init__(self):
        self.nodes = []          # (op, args)
        self.memo = {}

    def _mk(self, op, *args):
        key = (op,) + args
        if key not in self.memo:
            self.memo[key] = len(self.nodes)
            self.nodes.append(key)
        return self.memo[key]

    # signed reference = (sign, node)
    def inp(self, name):
        return (1, self._mk("in", name))

    def add(self, x, y):
        (sx, nx), (sy, ny) = x, y
        if sx == sy:
            a, b = sorted((nx, ny))
            return (sx, self._mk("add", a, b))
        if sx > 0:
            return (1, self._mk("sub", nx, ny))
        return (1, self._mk("sub", ny, nx))

    def sub(self, x, y):
        return self.add(x, (-y[0], y[1]))

    def mulc(self, c, x):
        if abs(c - 1) < EPS:
            return x
        if abs(c + 1) < EPS:
            return (-x[0], x[1])
        c = c * x[0]
        if c < 0:
            return (-1, self._mk("mul", repr(float(-c)), x[1]))
        return (1, self._mk("mul", repr(float(c)), x[1]))

    def fmac(self, c, x, y):
        """c*x + y"""
        cc = c * x[0] * y[0]
        return (y[0], self._mk("fma", repr(float(cc)), x[1], y[1]))


def cmul_const(d, z, w):
    """complex z times constant w=(wr,wi)."""
    (xr, xi), (wr, wi) = z, w
    if abs(wi) < EPS:
        return (d.mulc(wr, xr), d.mulc(wr, xi))
    if abs(wr) < EPS:
        return (d.mulc(-wi, xi), d.mulc(wi, xr))
    re = d.fmac(-wi, xi, d.mulc(wr, xr))
    im = d.fmac(wr, xi, d.mulc(wi, xr))
    return (re, im)


def cadd(d, a, b):
    return (d.add(a[0], b[0]), d.add(a[1], b[1]))


def csub(d, a, b):
    return (d.sub(a[0], b[0]), d.sub(a[1], b[1]))


def mul_neg_i(z):                       # z * (-i) = (im, -re)
    (sr, nr), (si, ni) = z
    return ((si, ni), (-sr, nr))


def mul_pos_i(z):                       # z * (+i) = (-im, re)
    (sr, nr), (si, ni) = z
    return ((-si, ni), (sr, nr))


def bf2(d, x):
    return [cadd(d, x[0], x[1]), csub(d, x[0], x[1])]


def bf4(d, x):
    t0, t1 = cadd(d, x[0], x[2]), csub(d, x[0], x[2])
    t2, t3 = cadd(d, x[1], x[3]), csub(d, x[1], x[3])
    return [cadd(d, t0, t2), cadd(d, t1, mul_neg_i(t3)), csub(d, t0, t2), cadd(d, t1, mul_pos_i(t3))]


def bf3(d, x):
    s = math.sin(2 * math.pi / 3)
    t1 = cadd(d, x[1], x[2])
    t2 = csub(d, x[1], x[2])
    y0 = cadd(d, x[0], t1)
    m1 = (d.fmac(-0.5, t1[0], x[0][0]), d.fmac(-0.5, t1[1], x[0][1]))
    # y1 = m1 - i s t2 ; y2 = m1 + i s t2
    y1 = (d.fmac(s, t2[1], m1[0]), d.fmac(-s, t2[0], m1[1]))
    y2 = (d.fmac(-s, t2[1], m1[0]), d.fmac(s, t2[0], m1[1]))
    return [y0, y1, y2]


def bf5(d, x):
    c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
    s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
    t1, t3 = cadd(d, x[1], x[4]), csub(d, x[1], x[4])
    t2, t4 = cadd(d, x[2], x[3]), csub(d, x[2], x[3])
    t5 = cadd(d, t1, t2)
    y0 = cadd(d, x[0], t5)
    ca, cb = (c1 + c2) / 2, (c1 - c2) / 2          # -1/4, sqrt(5)/4
    m1 = (d.fmac(ca, t5[0], x[0][0]), d.fmac(ca, t5[1], x[0][1]))
    dd = csub(d, t1, t2)
    sA = (d.fmac(cb, dd[0], m1[0]), d.fmac(cb, dd[1], m1[1]))
    sB = (d.fmac(-cb, dd[0], m1[0]), d.fmac(-cb, dd[1], m1[1]))
    # u = s1 t3 + s2 t4 ; v = s2 t3 - s1 t4   (to be multiplied by -i)
    u = (d.fmac(s1, t3[0], d.mulc(s2, t4[0])), d.fmac(s1, t3[1], d.mulc(s2, t4[1])))
    v = (d.fmac(s2, t3[0], d.mulc(-s1, t4[0])), d.fmac(s2, t3[1], d.mulc(-s1, t4[1])))
    y1, y4 = cadd(d, sA, mul_neg_i(u)), cadd(d, sA, mul_pos_i(u))
    y2, y3 = cadd(d, sB, mul_neg_i(v)), cadd(d, sB, mul_pos_i(v))
    return [y0, y1, y2, y3, y4]


PRIM = {2: bf2, 3: bf3, 4: bf4, 5: bf5}


def dft(d, x, factors):
    """x: list of n complex refs; factors: tuple of primitive radices, product n.
    First factor = first (DIF) stage over inputs strided by n/factors[0]."""
    n = len(x)
    if n == 1:
        return x
    n1 = factors[0]
    n2 = n // n1
    if n2 == 1:
        return PRIM[n1](d, x)
    # stage 1: for each j in [0,n2): n1-point DFT over x[j + n2*i]
    y = [[None] * n2 for _ in range(n1)]
    for j in range(n2):
        sub = PRIM[n1](d, [x[j + n2 * i] for i in range(n1)])
        for k1 in range(n1):
            ang = -2 * math.pi * ((j * k1) % n) / n
            w = (math.cos(ang), math.sin(ang))
            # snap exact values
            w = tuple(round(v) if abs(v - round(v)) < EPS else v for v in w)
            y[k1][j] = cmul_const(d, sub[k1], w)
    out = [None] * n
    for k1 in range(n1):
        z = dft(d, y[k1], factors[1:])
        for k2 in range(n2):
            out[k1 + n1 * k2] = z[k2]
    return out


def factorizations(n):
    def rec(m):
        if m == 1:
            yield ()
            return
        for p in (2, 3, 4, 5):
            if m % p == 0:
                for rest in rec(m // p):
                    yield (p,) + rest
    return sorted(set(rec(n)))


def live_nodes(d, outs):
    live, stack = set(), [n for o in outs for (_, n) in o]
    while stack:
        n = stack.pop()
        if n in live:
            continue
        live.add(n)
        node = d.nodes[n]
        if node[0] in ("add", "sub"):
            stack += [node[1], node[2]]
        elif node[0] == "mul":
            stack.append(node[2])
        elif node[0] == "fma":
            stack += [node[2], node[3]]
    return live


def build(n, factors):
    d = Dag()
    x = [(d.inp(f"re[{j}]"), d.inp(f"im[{j}]")) for j in range(n)]
    outs = dft(d, x, factors)
    live = live_nodes(d, outs)
    cost = sum(1 for i in live if d.nodes[i][0] != "in")
    return d, outs, live, cost


def fl(c):
    return f"{float(c)!r}f"


def emit(n, d, outs, live, cost, factors):
    """One codelet, generic over the value type V: float (one transform) or f2 (two independent transforms in the
    two halves of a 64-bit register pair -> packed FADD2 / FMUL2 / FFMA2 on sm_100a, see packed.cuh)."""
    L = [f"// dft{n}: factors {factors}, {cost} fp32 ops ({cost / n:.1f}/point)",
         f"template <class V> B2S_HD void dft{n}(V (&re)[{n}], V (&im)[{n}]) {{"]
    name = {}
    for i, node in enumerate(d.nodes):
        if i not in live:
            continue
        op = node[0]
        if op == "in":
            name[i] = node[1]
            continue
        name[i] = f"t{i}"
        if op == "add":
            e = f"vadd({name[node[1]]}, {name[node[2]]})"
        elif op == "sub":
            e = f"vsub({name[node[1]]}, {name[node[2]]})"
        elif op == "mul":
            e = f"vmul({fl(node[1])}, {name[node[2]]})"
        else:
            e = f"vfma({fl(node[1])}, {name[node[2]]}, {name[node[3]]})"
        L.append(f"  const V t{i} = {e};")
    for k, ((sr, nr), (si, ni)) in enumerate(outs):
        L.append(f"  const V o{k}r = {'vneg(' + name[nr] + ')' if sr < 0 else name[nr]};"
                 f" const V o{k}i = {'vneg(' + name[ni] + ')' if si < 0 else name[ni]};")
    for k in range(n):
        L.append(f"  re[{k}] = o{k}r; im[{k}] = o{k}i;")
    L.append("}")
    return "\n".join(L)


def main():
    parts = ["// GENERATED by tools/gen_codelets.py — do not edit.",
             "// Forward DFT codelets (sign -1), natural order, literal twiddles.",
             "#pragma once",
             "#ifndef B2S_HD",
             "#if defined(__CUDACC__)",
             "#define B2S_HD __host__ __device__ __forceinline__",
             "#else",
             "#define B2S_HD inline",
             "#endif",
             "#endif",
             "#include <math.h>",
             '#include "packed.cuh"',
             "namespace b2s {", ""]
    summary = []
    for n in SIZES:
        best = None
        for f in factorizations(n):
            for perm in set(itertools.permutations(f)):
                cand = build(n, perm)
                if best is None or cand[3] < best[0][3]:
                    best = (cand, perm)
        (d, outs, live, cost), perm = best
        parts.append(emit(n, d, outs, live, cost, perm))
        parts.append("")
        summary.append((n, perm, cost))
    parts.append("template <int N> struct Dft;")
    for n in SIZES:
        parts.append(f"template <> struct Dft<{n}> {{ template <class V> static B2S_HD void run(V (&re)[{n}], V (&im)[{n}]) {{ dft{n}(re, im); }} }};")
    parts.append("template <> struct Dft<1> { template <class V> static B2S_HD void run(V (&)[1], V (&)[1]) {} };")
    parts.append("}  // namespace b2s")
    OUT.parent.mkdir(parents=True, exist_ok=True)
    OUT.write_text("\n".join(parts) + "\n")
    for n, perm, cost in summary:
        print(f"dft{n:<3d} {str(perm):<16s} {cost:5d} ops  {cost / n:5.1f}/pt")


if __name__ == "__main__":
    main()
