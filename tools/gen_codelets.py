#!/usr/bin/env python3
"""Generator of the straight-line fp64 FFT codelets in dynpy_b200/csrc/codelets.cuh.

The row/column kernels are bound by the fp64 issue rate (64 lanes/clk/SM on B200), so the
codelets are written to minimise the number of fp64 INSTRUCTIONS, not flops:

  * every real value carries a compile-time "pending" coefficient; an addition of two values
    with different coefficients is one FMA (a + (cb/ca) b, pending ca), so multiplications by
    real constants are free, multiplications by -i / i are register renames, and a
    multiplication by a general complex constant costs 2 FMAs instead of 4 instructions;
  * structurally zero inputs (zero padding of the Wiener-Khinchin transform) are propagated
    symbolically, which prunes the first stages of the column transforms;
  * power-of-two sizes are radix-2 DIF recursions, size 5 is the symmetric-pair form of the
    prime-length DFT, 25 = 5 x 5 Cooley-Tukey with constant twiddles.

Every codelet takes natural-order inputs v[0..n) (cd = double2) and leaves X[k] in v[k].
`python tools/gen_codelets.py` rewrites codelets.cuh and prints the instruction counts;
tests/test_codelets.py compiles the header for the host and checks every codelet against
numpy.fft.
"""
import cmath
import math
import os
import sys

EPS = 1e-15


class Term:
    """coef * reg  (reg None == structural zero)."""
    __slots__ = ("reg", "coef")

    def __init__(self, reg, coef=1.0):
        self.reg = reg
        self.coef = coef if reg is not None else 0.0

    def scaled(self, c):
        if self.reg is None or abs(c) < EPS:
            return Term(None)
        return Term(self.reg, self.coef * c)


ZERO = Term(None)


def fconst(x):
    r = repr(float(x))
    if "e" not in r and "." not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return r


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0
        self.ops = {"add": 0, "fma": 0, "mul": 0}

    def tmp(self):
        self.n += 1
        return "t%d" % self.n

    def lin(self, a, b):
        """a + b for Terms (signs live in the coefficients); one instruction."""
        if a.reg is None:
            return b
        if b.reg is None:
            return a
        # choose which coefficient stays pending: prefer +1, then -1, then a's
        keep_a = True
        if abs(abs(a.coef) - 1.0) > EPS and abs(abs(b.coef) - 1.0) < EPS:
            keep_a = False
        elif abs(a.coef + 1.0) < EPS and abs(b.coef - 1.0) < EPS:
            keep_a = False
        if not keep_a:
            a, b = b, a
        ratio = b.coef / a.coef
        t = self.tmp()
        if abs(ratio - 1.0) < EPS:
            self.lines.append("const double %s = %s + %s;" % (t, a.reg, b.reg))
            self.ops["add"] += 1
        elif abs(ratio + 1.0) < EPS:
            self.lines.append("const double %s = %s - %s;" % (t, a.reg, b.reg))
            self.ops["add"] += 1
        else:
            self.lines.append("const double %s = fma(%s, %s, %s);" % (t, fconst(ratio), b.reg, a.reg))
            self.ops["fma"] += 1
        return Term(t, a.coef)

    def materialise(self, a):
        """register holding the value of the Term (coef applied)."""
        if a.reg is None:
            return "0.0"
        if abs(a.coef - 1.0) < EPS:
            return a.reg
        t = self.tmp()
        if abs(a.coef + 1.0) < EPS:
            self.lines.append("const double %s = -%s;" % (t, a.reg))
        else:
            self.lines.append("const double %s = %s * %s;" % (t, fconst(a.coef), a.reg))
        self.ops["mul"] += 1
        return t


class C:
    """complex value = (Term re, Term im)."""
    __slots__ = ("re", "im")

    def __init__(self, re, im):
        self.re, self.im = re, im


def cadd(E, x, y):
    return C(E.lin(x.re, y.re), E.lin(x.im, y.im))


def csub(E, x, y):
    return C(E.lin(x.re, y.re.scaled(-1.0)), E.lin(x.im, y.im.scaled(-1.0)))


def cmulc(E, x, w):
    """x * w for a compile-time complex constant w."""
    wr, wi = w.real, w.imag
    if abs(wr) < EPS:
        wr = 0.0
    if abs(wi) < EPS:
        wi = 0.0
    re = E.lin(x.re.scaled(wr), x.im.scaled(-wi))
    im = E.lin(x.re.scaled(wi), x.im.scaled(wr))
    return C(re, im)


def cscale(x, c):
    return C(x.re.scaled(c), x.im.scaled(c))


def root(n, k):
    """exp(-2 pi i k / n) with exact values on the axes / diagonals."""
    k %= n
    g = math.gcd(k, n)
    k //= g
    n //= g
    if n == 1:
        return 1 + 0j
    if n == 2:
        return -1 + 0j
    if n == 4:
        return {1: -1j, 3: 1j}[k]
    if n == 8:
        r = math.sqrt(0.5)
        return {1: complex(r, -r), 3: complex(-r, -r), 5: complex(-r, r), 7: complex(r, r)}[k]
    return cmath.exp(-2j * cmath.pi * k / n)


def dft_prime(E, x):
    """odd-prime DFT through the symmetric pairs t_j = x_j + x_{p-j}, s_j = x_j - x_{p-j}."""
    p = len(x)
    h = (p - 1) // 2
    t = [cadd(E, x[j], x[p - j]) for j in range(1, h + 1)]
    s = [csub(E, x[j], x[p - j]) for j in range(1, h + 1)]
    X = [None] * p
    acc = x[0]
    for j in range(h):
        acc = cadd(E, acc, t[j])
    X[0] = acc
    for k in range(1, h + 1):
        A = x[0]
        for j in range(1, h + 1):
            A = cadd(E, A, cscale(t[j - 1], math.cos(2 * math.pi * j * k / p)))
        B = None
        for j in range(1, h + 1):
            term = cscale(s[j - 1], math.sin(2 * math.pi * j * k / p))
            B = term if B is None else cadd(E, B, term)
        # X_k = A - i B ,  X_{p-k} = A + i B ;   -iB = (B.im, -B.re)
        mB = C(B.im, B.re.scaled(-1.0))
        X[k] = cadd(E, A, mB)
        X[p - k] = csub(E, A, mB)
    return X


def fft(E, x):
    n = len(x)
    if n == 1:
        return x
    if n % 2 == 0:
        half = n // 2
        a = [cadd(E, x[j], x[j + half]) for j in range(half)]
        b = [cmulc(E, csub(E, x[j], x[j + half]), root(n, j)) for j in range(half)]
        Ev, Od = fft(E, a), fft(E, b)
        X = [None] * n
        for k in range(half):
            X[2 * k], X[2 * k + 1] = Ev[k], Od[k]
        return X
    if n in (3, 5, 7):
        return dft_prime(E, x)
    # odd composite: Cooley-Tukey n = r * m with the smallest prime factor r
    r = next(q for q in (3, 5, 7) if n % q == 0)
    m = n // r
    # n1 = m*q + j  (q < r, j < m); k = k1 + r*k2 : first r-point DFTs over q, twiddle W_n^{j k1}, then m-point
    cols = []
    for j in range(m):
        y = fft(E, [x[m * q + j] for q in range(r)])
        cols.append([cmulc(E, y[k1], root(n, j * k1)) for k1 in range(r)])
    X = [None] * n
    for k1 in range(r):
        z = fft(E, [cols[j][k1] for j in range(m)])
        for k2 in range(m):
            X[k1 + r * k2] = z[k2]
    return X


def gen_codelet(name, n, nonzero=None, pre=None, doc=""):
    """void name(cd (&v)[n]) : v <- DFT_n(v); inputs >= nonzero are structurally zero (not read);
    pre[j] (optional) is a constant each input j is multiplied by first."""
    E = Emitter()
    nz = n if nonzero is None else nonzero
    x = []
    for j in range(n):
        if j < nz:
            c = C(Term("v[%d].x" % j), Term("v[%d].y" % j))
            if pre is not None:
                c = cmulc(E, c, pre[j])
            x.append(c)
        else:
            x.append(C(ZERO, ZERO))
    X = fft(E, x)
    outs = []
    for k in range(n):
        outs.append((E.materialise(X[k].re), E.materialise(X[k].im)))
    body = ["    " + l for l in E.lines]
    for k, (r, i) in enumerate(outs):
        body.append("    v[%d] = make_double2(%s, %s);" % (k, r, i))
    total = sum(E.ops.values())
    head = "// %s: %d-point DFT%s — %d fp64 instructions (%d add, %d fma, %d mul)%s" % (
        name, n, "" if nz == n else " of %d leading non-zero inputs" % nz, total, E.ops["add"], E.ops["fma"],
        E.ops["mul"], (" — " + doc) if doc else "")
    src = "%s\nDYNPY_HD void %s(cd (&v)[%d]) {\n%s\n}\n" % (head, name, n, "\n".join(body))
    return src, total, E.ops


HEADER = """// GENERATED by tools/gen_codelets.py — do not edit by hand.
// Straight-line fp64 DFT codelets (natural-order in, natural-order out, in place in v[]).
// Instruction-count minimised: constant factors ride along as pending FMA coefficients.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#ifdef __CUDACC__
#define DYNPY_HD __host__ __device__ __forceinline__
#else
#define DYNPY_HD static inline
#endif
namespace dynpy {
typedef double2 cd;
"""


def main():
    out = [HEADER]
    report = []
    specs = [
        ("fft2", 2, None, None, ""),
        ("fft4", 4, None, None, ""),
        ("fft8", 8, None, None, ""),
        ("fft16", 16, None, None, ""),
        ("fft2_h1", 2, 1, None, "zero-padded half"),
        ("fft4_h2", 4, 2, None, "zero-padded half"),
        ("fft8_h4", 8, 4, None, "zero-padded half"),
        ("fft16_h8", 16, 8, None, "zero-padded half"),
        ("fft32", 32, None, None, ""),
        ("fft8_odd", 8, None, [root(16, j) for j in range(8)], "odd outputs of a 16-point DFT of 8 leading non-zero inputs"),
        ("fft16_odd", 16, None, [root(32, j) for j in range(16)], "odd outputs of a 32-point DFT of 16 leading non-zero inputs"),
        ("fft32_odd", 32, None, [root(64, j) for j in range(32)], "odd outputs of a 64-point DFT of 32 leading non-zero inputs"),
        ("fft10", 10, None, None, "2 x 5"),
        ("fft10_odd", 10, None, [root(20, j) for j in range(10)], "odd outputs of a 20-point DFT of 10 leading non-zero inputs"),
        ("fft12", 12, None, None, "4 x 3"),
        ("fft12_odd", 12, None, [root(24, j) for j in range(12)], "odd outputs of a 24-point DFT of 12 leading non-zero inputs"),
        ("fft20", 20, None, None, "4 x 5"),
        ("fft20_odd", 20, None, [root(40, j) for j in range(20)], "odd outputs of a 40-point DFT of 20 leading non-zero inputs"),
        ("fft5", 5, None, None, ""),
        ("fft25", 25, None, None, "5 x 5"),
        ("fft25_odd", 25, None, [root(50, j) for j in range(25)],
         "odd outputs of a 50-point DFT of 25 leading non-zero inputs: DFT_25(v[j] W_50^j)"),
    ]
    for name, n, nz, pre, doc in specs:
        src, total, ops = gen_codelet(name, n, nz, pre, doc)
        out.append(src)
        report.append((name, n, total, ops))
    out.append("}  // namespace dynpy\n")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "dynpy_b200", "csrc", "codelets.cuh")
    with open(path, "w") as f:
        f.write("\n".join(out))
    for name, n, total, ops in report:
        print("%-10s n=%2d  %4d instr  (%.2f per point)  %s" % (name, n, total, total / n, ops))


if __name__ == "__main__":
    main()
