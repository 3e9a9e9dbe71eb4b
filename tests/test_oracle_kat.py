"""Pins the oracle against the reference's own known answers (no GPU needed).

  - the 102 shipped Float64 tables of the b-series and of a3 / a4 / a5 / a7 (src/precomputed_kernels/kernel_*.jl),
    antiderivative tables (_i1) and their D1 / D2 companions included - bit equality
  - the integer ghost matrices (src/convolution_coefs/polynomial_boundary.jl:29-179) - exact
  - the exact-rational kernel polynomials regenerate the tables (independent Fraction path)
"""
import json
import os
import re
from fractions import Fraction

import numpy as np
import pytest


def _golden_tables(golden_dir):
    return np.load(os.path.join(golden_dir, "shipped_tables.npz"))


def test_oracle_tables_bit_equal_shipped_literals(oracle, golden_dir):
    g = _golden_tables(golden_dir)
    n = 0
    for name in g.files:
        m = re.match(r"KERNEL_(D1_|D2_|)PRE_([ab]\d+)_(d|i)(\d+)$", name)
        base = int(m.group(4)) if m.group(3) == "d" else -int(m.group(4))  # _i1: the antiderivative family
        kernel, d = m.group(2), base + {"": 0, "D1_": 1, "D2_": 2}[m.group(1)]
        t = oracle.table_i1(kernel) if d == -1 else oracle.table(kernel, d)
        lit = g[name]
        assert t.shape == lit.shape
        same = (t.view(np.uint64) == lit.view(np.uint64)) | ((t == 0) & (lit == 0))
        assert same.all(), name
        n += 1
    assert n == 102


def test_d1_companion_is_next_derivative_table(golden_dir):
    # KERNEL_D1_PRE_<k>_d<d> is the same matrix as KERNEL_PRE_<k>_d<d+1> (precompute_kernel_and_range.jl:80-99)
    g = _golden_tables(golden_dir)
    for k, dmax in (("b5", 1), ("b13", 4)):
        for d in range(dmax + 1):
            assert np.array_equal(g[f"KERNEL_D1_PRE_{k}_d{d}"], g[f"KERNEL_PRE_{k}_d{d + 1}"])


@pytest.mark.parametrize("kernel,d,rows", [("b5", 0, range(0, 101, 7)), ("b5", 2, range(0, 101, 9)),
                                           ("b13", 0, (0, 37, 100)), ("b13", 5, (1, 50, 99)),
                                           ("a3", 0, range(0, 101, 11)), ("a4", 0, range(0, 101, 13)),
                                           ("a5", 0, (0, 33, 100)), ("a7", 0, (2, 51, 99))])
def test_rational_polynomials_regenerate_tables(kernel, d, rows, golden_dir):
    root = os.path.dirname(os.path.dirname(golden_dir))
    polys = json.load(open(os.path.join(root, "convolutioninterpolations.jl_b200", "kernel_polys.json")))[kernel]
    eqs = polys["eqs"]
    pieces = [[Fraction(c) for c in p] for p in polys["pieces"]]
    lit = _golden_tables(golden_dir)[f"KERNEL_PRE_{kernel}_d{d}"]
    for r in rows:
        for c in range(2 * eqs):
            s = Fraction(r, 100) - eqs + c  # precompute_kernel_and_range.jl:150-157
            sa = abs(s)
            if sa >= eqs:
                v = Fraction(0)
            else:
                coef = pieces[int(sa)]
                dc = []
                for k in range(d, len(coef)):
                    f = Fraction(1)
                    for j in range(d):
                        f *= k - j
                    dc.append(f * coef[k])
                v = Fraction(0)
                for ck in reversed(dc):
                    v = v * sa + ck  # horner.jl:44-52
                if d % 2 == 1:
                    v *= (s > 0) - (s < 0)  # convolution_kernel_b5.jl:79-88
            assert v.numerator / v.denominator == lit[r, c]


def test_ghost_matrices_match_reference_literals(oracle, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "ghost_matrices.json")))
    for kernel in ("b5", "b7", "b9", "b11", "b13", "a3", "a4", "a5", "a7"):
        lit = np.array(g[f"poly_{kernel}"])
        assert np.array_equal(oracle.ghost_matrix("poly", kernel, lit.shape[0]), lit)
    assert np.array_equal(oracle.ghost_matrix("linear", "b13", 9), np.array(g["linear"]))
    assert np.array_equal(oracle.ghost_matrix("quadratic", "b13", 9), np.array(g["quadratic"]))
    # every row sums to 1 (constant reproduction) - SURVEY.md 3.4
    for key, rows in g.items():
        assert all(abs(sum(r) - 1.0) == 0 for r in rows), key


def test_table_properties(oracle):
    # partition of unity of the interpolating kernel: each table row of K sums to 1, of K' to 0
    for k in ("b5", "b13", "a3", "a4", "a5", "a7"):
        t0, t1 = oracle.table(k, 0), oracle.table(k, 1)
        assert np.abs(t0.sum(axis=1) - 1).max() < 1e-14
        assert np.abs(t1.sum(axis=1)).max() < 1e-12
        e = t0.shape[1] // 2
        assert t0[0, e] == 1.0 and t0[100, e - 1] == 1.0  # K(0) = 1 at both ends of the table


def test_antiderivative_tables(oracle):
    # K~ is odd, saturates at -+1/2 at the ends of the support, and its table rows telescope: K~(s + 1 cell) - K~(s)
    # summed over the columns = K~(eqs) - K~(-eqs) = 1 (precomputed_kernel_tables.jl:93-96, KERNEL_PRE_<k>_i1)
    for k in ("b5", "b7", "b9", "b11", "b13", "a3", "a4", "a5", "a7"):
        t = oracle.table_i1(k)
        assert t[0, 0] == -0.5 and t[100, -1] == 0.5
        assert np.array_equal(t[100, :-1], t[0, 1:])          # row 100 of column c is row 0 of column c + 1
        assert np.abs(t + t[::-1, ::-1]).max() < 1e-16         # odd
        assert np.abs((t[100] - t[0]).sum() - 1.0) < 1e-15
