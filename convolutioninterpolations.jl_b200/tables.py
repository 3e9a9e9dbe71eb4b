"""Generated kernel tables for :linear-subgrid axes.

The reference raises the table length of every :linear axis to max(precompute, 10_000) rows
(fast_convolution_interpolation_constructor.jl:117-118, convolution_resample.jl:228-229) and fills them with
    kernel_pre[r, c] = K^(d)((r-1)/(n_pre-1) - eqs + c - 1),  exact rational arithmetic, rounded once
(precompute_kernel_and_range.jl:40-157).  A Julia caller passes its own itp.kernel_pre through the C-ABI; this module is
what the Python host mirror uses instead: the same numbers from the exact rational kernel polynomials
(kernel_polys.json, written by tools/gen_tables.py), with integer Horner steps and one correctly rounded division per
entry.  n_pre = 101 reproduces the shipped tables bit for bit (tests/test_host_logic.py)."""
import functools
import json
import math
import os
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LINEAR_ROWS = 10_000


@functools.lru_cache(maxsize=None)
def _polys():
    return json.load(open(os.path.join(_HERE, "kernel_polys.json")))


@functools.lru_cache(maxsize=None)
def _integer_pieces(kernel, d):
    """per piece: integer coefficients A_k and the common denominator L of the d-th derivative polynomial in |s|"""
    info = _polys()[kernel]
    out = []
    for piece in info["pieces"]:
        c = [Fraction(x) for x in piece]
        dc = []
        for k in range(d, len(c)):
            f = Fraction(1)
            for j in range(d):
                f *= k - j
            dc.append(f * c[k])
        if not dc:
            dc = [Fraction(0)]
        L = 1
        for x in dc:
            L = L * x.denominator // math.gcd(L, x.denominator)
        out.append(([int(x * L) for x in dc], L))
    return info["eqs"], out


@functools.lru_cache(maxsize=None)
def _integer_pieces_antiderivative(kernel):
    """per piece p (|s| in [p, p+1)): integer coefficients and common denominator of K~(|s|) = int_0^|s| K(u) du as ONE
    polynomial in |s| - the integration constant carries the integral over the pieces below (the reference's
    b5_coefs_i1 ... are these rationals: convolution_kernel_b5.jl:52-58)"""
    info = _polys()[kernel]
    out = []
    below = Fraction(0)
    for p, piece in enumerate(info["pieces"]):
        c = [Fraction(x) for x in piece]
        a = [Fraction(0)] + [c[k] / (k + 1) for k in range(len(c))]           # antiderivative without constant
        at = lambda s, a=a: sum(ak * s ** k for k, ak in enumerate(a))
        a[0] = below - at(Fraction(p))                                        # continuity at |s| = p
        below += at(Fraction(p + 1)) - at(Fraction(p))
        L = 1
        for x in a:
            L = L * x.denominator // math.gcd(L, x.denominator)
        out.append(([int(x * L) for x in a], L))
    return info["eqs"], out


@functools.lru_cache(maxsize=32)
def generated_table(kernel, d, n_pre=LINEAR_ROWS):
    """(n_pre, 2*eqs) C-ordered Float64 table of K^(d) (d = -1: the antiderivative K~, odd, saturating at +-1/2);
    cached per process"""
    if d == -1:
        eqs, pieces = _integer_pieces_antiderivative(kernel)
        Q = n_pre - 1
        t = np.zeros((n_pre, 2 * eqs))
        for c in range(2 * eqs):
            for r in range(n_pre):
                m = r + (c - eqs) * Q
                M = -m if m < 0 else m
                if M >= eqs * Q:
                    t[r, c] = 0.5 if m > 0 else -0.5
                    continue
                A, L = pieces[M // Q]
                n = len(A) - 1
                acc, qp = A[n], 1
                for k in range(n - 1, -1, -1):
                    qp *= Q
                    acc = acc * M + A[k] * qp
                if m < 0:
                    acc = -acc
                t[r, c] = acc / (L * qp)
        t.setflags(write=False)
        return t
    eqs, pieces = _integer_pieces(kernel, d)
    Q = n_pre - 1
    t = np.zeros((n_pre, 2 * eqs))
    for c in range(2 * eqs):
        for r in range(n_pre):
            m = r + (c - eqs) * Q          # s = m / Q
            M = -m if m < 0 else m
            if M >= eqs * Q:
                continue                   # outside the support: 0
            A, L = pieces[M // Q]
            n = len(A) - 1
            acc, qp = A[n], 1
            for k in range(n - 1, -1, -1):
                qp *= Q
                acc = acc * M + A[k] * qp
            if d % 2 == 1 and m < 0:
                acc = -acc                 # odd derivatives are odd functions (convolution_kernel_b5.jl:79-92)
            t[r, c] = acc / (L * qp)       # int / int: correctly rounded
    t.setflags(write=False)
    return t
