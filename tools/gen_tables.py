#!/usr/bin/env python3
"""Regenerate the 101-row kernel tables of the reference from exact rationals.

Run in the authoring container only (needs /root/reference, read-only):

    python tools/gen_tables.py

What it does
  1. reads the exact rational piecewise-polynomial coefficients of the b-series
     and a-series (a3, a4, a5, a7) kernels (reference: src/convolution_kernels/convolution_kernel_b*.jl, the
     `<k>_coefs = Dict(:eq1 => [...], ...)` literal; derivative dictionaries are
     NOT read, they are re-derived as k*(k-1)*... * c_k),
  2. regenerates every 101 x 2*eqs table with `fractions.Fraction` following the
     generator rule of the reference
     (src/convolution_fast_interpolation/precompute_kernel_and_range.jl:150-157:
      table[r, c] = K^(d)((r-1)/100 - eqs + c - 1), exact, then rounded once),
     using the Horner-with-skip and odd-derivative sign rules
     (src/convolution_kernels/horner.jl:44-52, convolution_kernel_b5.jl:79-92),
  3. parses the shipped Float64 literals
     (src/precomputed_kernels/kernel_b*.jl) and asserts BIT equality
     (this is the known-answer test that pins table layout + kernel maths),
  4. writes
       tests/golden/shipped_tables.npz           the shipped literals (golden KAT)
       tests/golden/ghost_matrices.json          the integer ghost matrices (golden KAT)
       convolutioninterpolations.jl_b200/csrc/cib200_tables.inc   C hex-float arrays

Nothing here is imported at run time by the product, the tests or the bench.
"""
import json
import os
import re
import sys
from fractions import Fraction

import numpy as np

REF = "/root/reference/src"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel -> (eqs, highest analytic derivative that has a coefficient dict)
KERNELS = {"b5": (5, 3), "b7": (6, 4), "b9": (7, 5), "b11": (8, 6), "b13": (9, 6),
           # a-series (Keys-type cubic pieces): only K and K' are shipped / needed (derivative 0, cubic subgrid)
           "a3": (2, 1), "a4": (3, 1), "a5": (3, 1), "a7": (4, 1)}
N_PRE = 101


def _parse_rational_expr(expr):
    """Julia `a//b`, `big(a)//big(b)`, products thereof -> Fraction (left-assoc)."""
    expr = expr.strip().replace("big(", "(")
    expr = re.sub(r"(?<![\w.])(-?\d+)", r"Fraction(\1)", expr)
    expr = expr.replace("//", "/")
    # `a` is the free parameter of the :a5 kernel (convolution_kernel_a5.jl:8: const a = 3//64)
    return eval(expr, {"Fraction": Fraction, "a": Fraction(3, 64)})  # noqa: S307 - trusted local literal


def read_base_coefs(kernel):
    path = f"{REF}/convolution_kernels/convolution_kernel_{kernel}.jl"
    src = open(path).read()
    m = re.search(rf"const {kernel}_coefs = Dict\((.*?)\n\)", src, re.S)
    assert m, path
    body = m.group(1)
    out = {}
    for em in re.finditer(r":eq(\d+)\s*=>\s*\[(.*?)\]", body, re.S):
        idx = int(em.group(1))
        items = _split_top_level(em.group(2))
        out[idx] = [_parse_rational_expr(x) for x in items]
    eqs = KERNELS[kernel][0]
    assert sorted(out) == list(range(1, eqs + 1)), (kernel, sorted(out))
    return [out[i] for i in range(1, eqs + 1)]


def _split_top_level(s):
    items, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            items.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        items.append(cur)
    return items


def deriv_coefs(c, d):
    """d-th derivative of sum c[k] s^k, as coefficient list of s^(k-d), k>=d."""
    out = []
    for k in range(d, len(c)):
        f = Fraction(1)
        for j in range(d):
            f *= (k - j)
        out.append(f * c[k])
    return out


def kernel_value(polys, eqs, d, s):
    """K^(d)(s) for exact rational s (reference: convolution_kernel_b5.jl:50-92)."""
    sa = abs(s)
    if sa >= eqs:
        return Fraction(0)
    piece = int(sa)  # pieces are [0,1), [1,2), ...
    c = deriv_coefs(polys[piece], d)
    acc = Fraction(0)
    for ck in reversed(c):
        acc = acc * sa + ck
    if d % 2 == 1:
        sign = (s > 0) - (s < 0)
        acc *= sign
    return acc


def kernel_antiderivative(polys, eqs, s):
    """K~(s) = int_0^s K(u) du (odd, saturating at +-1/2 for |s| >= eqs): the `_i1` tables
    (precomputed_kernels/kernel_<k>.jl KERNEL_PRE_<k>_i1; selector precomputed_kernel_tables.jl:93-96)."""
    sa = min(abs(s), Fraction(eqs))
    acc = Fraction(0)
    for piece in range(eqs):
        lo, hi = Fraction(piece), min(sa, Fraction(piece + 1))
        if hi <= lo:
            break
        acc += sum(ck * (hi ** (k + 1) - lo ** (k + 1)) / (k + 1) for k, ck in enumerate(polys[piece]))
    return acc * ((s > 0) - (s < 0))


def make_antiderivative_table(polys, eqs):
    t = np.zeros((N_PRE, 2 * eqs))
    for r in range(N_PRE):
        for c in range(2 * eqs):
            v = kernel_antiderivative(polys, eqs, Fraction(r, N_PRE - 1) - eqs + c)
            t[r, c] = v.numerator / v.denominator
    return t


def make_table(polys, eqs, d):
    t = np.zeros((N_PRE, 2 * eqs))
    for r in range(N_PRE):
        for c in range(2 * eqs):
            s = Fraction(r, N_PRE - 1) - eqs + c  # 0-based r, c
            v = kernel_value(polys, eqs, d, s)
            t[r, c] = v.numerator / v.denominator  # correctly rounded true division
    return t


def read_shipped(kernel):
    path = f"{REF}/precomputed_kernels/kernel_{kernel}.jl"
    src = open(path).read()
    out = {}
    for m in re.finditer(r"const (KERNEL_\w+) = Float64\[(.*?)\n\]", src, re.S):
        rows = [r.strip() for r in m.group(2).split(";") if r.strip()]
        if not rows:
            continue
        out[m.group(1)] = np.array([[float(x) for x in r.split()] for r in rows])
    return out


def read_ghost_matrices():
    """integer ghost matrices: convolution_coefs/polynomial_boundary.jl:29-131,157-179"""
    src = open(f"{REF}/convolution_coefs/polynomial_boundary.jl").read()
    out = {}

    def rows_of(body):
        rows = []
        for line in body.splitlines():
            line = line.split("#")[0].strip().rstrip(";").strip()
            if line:
                rows.append([float(x) for x in line.split()])
        return rows
    for m in re.finditer(r":(\w+)\s*=>\s*\[(.*?)\]", src, re.S):
        out["poly_" + m.group(1)] = rows_of(m.group(2))
    for name, key in (("LINEAR_GHOST_MATRIX", "linear"), ("QUADRATIC_GHOST_MATRIX", "quadratic")):
        m = re.search(rf"const {name} = \[(.*?)\n\]", src, re.S)
        out[key] = rows_of(m.group(1))
    return out


def c_array(name, arr):
    flat = arr.reshape(-1)
    lines = [f"static const double {name}[{flat.size}] = {{"]
    for i in range(0, flat.size, 4):
        lines.append("  " + ", ".join(float(v).hex() for v in flat[i:i + 4]) + ",")
    lines.append("};")
    return "\n".join(lines)


def main():
    polys_json = {}
    golden = {}
    inc = ["/* GENERATED by tools/gen_tables.py - do not edit.",
           " * K^(d) tables, 101 rows x 2*eqs columns, row-major (and the antiderivative tables K~ = _i1):",
           " *   T[r][c] = K^(d)(r/100 - eqs + c), exact rational arithmetic rounded once.",
           " * Bit-identical to the Float64 literals shipped by the reference",
           " * (src/precomputed_kernels/kernel_<k>.jl), asserted at generation time. */"]
    n_checked = 0
    index_rows = []
    for kernel, (eqs, dmax) in KERNELS.items():
        polys = read_base_coefs(kernel)
        polys_json[kernel] = {
            "eqs": eqs, "max_derivative": dmax,
            "pieces": [[f"{c.numerator}/{c.denominator}" for c in p] for p in polys]}
        shipped = read_shipped(kernel)
        tabs = {d: make_table(polys, eqs, d) for d in range(dmax + 1)}
        tabs[-1] = make_antiderivative_table(polys, eqs)
        for name, lit in shipped.items():
            m = re.match(rf"KERNEL_(D1_|D2_|)PRE_{kernel}_(d|i)(\d+)$", name)
            if not m:
                continue
            # antiderivative family: K~ = _i1, its D1 / D2 companions are K and K'
            base = int(m.group(3)) if m.group(2) == "d" else -int(m.group(3))
            d = base + {"": 0, "D1_": 1, "D2_": 2}[m.group(1)]
            if d not in tabs:
                continue  # a-series K'' companions are not needed (integral directions use the cubic subgrid)
            assert lit.shape == tabs[d].shape, (name, lit.shape)
            same = lit.view(np.uint64) == tabs[d].view(np.uint64)
            # +0.0 vs -0.0 both legal for exact zeros of odd derivatives
            zero_both = (lit == 0) & (tabs[d] == 0)
            bad = ~(same | zero_both)
            assert not bad.any(), (name, int(bad.sum()))
            golden[name] = lit
            n_checked += lit.size
        inc.append(c_array(f"CIB_TAB_{kernel}_i1", tabs[-1]))
        for d in range(dmax + 1):
            cname = f"CIB_TAB_{kernel}_d{d}"
            inc.append(c_array(cname, tabs[d]))
            index_rows.append((kernel, eqs, d, cname))
    inc.append("")
    text = "\n".join(inc) + "\n"
    # ONE copy of each generated artefact: the library compiles the .inc, the oracle's Makefile includes the same
    # file, and the host mirror (tables.py) and the oracle both read the one kernel_polys.json
    p = os.path.join(ROOT, "convolutioninterpolations.jl_b200/csrc/cib200_tables.inc")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    open(p, "w").write(text)
    json.dump(polys_json, open(os.path.join(ROOT, "convolutioninterpolations.jl_b200/kernel_polys.json"), "w"), indent=1)
    os.makedirs(os.path.join(ROOT, "tests/golden"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "tests/golden/shipped_tables.npz"), **golden)
    json.dump(read_ghost_matrices(), open(os.path.join(ROOT, "tests/golden/ghost_matrices.json"), "w"), indent=1)
    print(f"bit-checked {n_checked} shipped doubles in {len(golden)} tables; "
          f"emitted {len(index_rows)} K^(d) tables")


if __name__ == "__main__":
    sys.exit(main())
