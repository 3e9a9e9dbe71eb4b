"""Second oracle: the reference's DIRECT evaluators (convolution_kernel_interpolation_eager_{1d,2d,3d}.jl:
kernel polynomials by Horner, no tables, no Hermite blending) restated in oracle.direct_eval, and the reference's own
assertion that the fast path agrees with them (test/test_constructors.jl:4-49, atol 1e-6) - here against the fast
restatement (CPU) and, in tests/test_gpu_eval.py, against the CUDA library."""
import numpy as np
import pytest

from oracle import oracle as O

KERNELS = ("a3", "a4", "a5", "a7", "b5", "b7", "b9", "b11", "b13")


@pytest.mark.parametrize("kernel", KERNELS)
def test_low_level_constructors_fast_agrees_with_direct(kernel):
    """test/test_constructors.jl:4-49 (a0/a1 and the lazy variant are outside the path)"""
    x = np.linspace(0.0, 2 * np.pi, 30)
    vs = np.sin(x)
    itp = O.Interpolant((x,), vs, kernel=kernel)
    direct = O.direct_eval(itp, np.array([1.0]))[0]
    fast = itp(1.0)
    assert abs(direct - np.sin(1.0)) <= 0.1      # :14
    assert abs(fast - np.sin(1.0)) <= 0.1        # :18
    assert abs(direct - fast) <= 1e-6            # :21
    etp = O.Interpolant((x,), vs, kernel=kernel, extrap="line")
    assert abs(etp(1.0) - np.sin(1.0)) <= 0.1    # :25
    assert np.isfinite(etp(-0.5))                # :26
    with pytest.raises(O.OracleError):           # :29
        O.Interpolant((x,), vs, kernel=kernel, extrap="throw")(-0.5)
    # 2-D (:36-47)
    v2 = np.sin(x)[:, None] * np.cos(x)[None, :]
    itp2 = O.Interpolant((x, x), v2, kernel=kernel, bc="detect", derivative=0, subgrid="cubic")
    d2 = O.direct_eval(itp2, np.array([1.0]), np.array([1.0]))[0]
    assert abs(d2 - itp2(1.0, 1.0)) <= 1e-6
    assert abs(d2 - np.sin(1.0) * np.cos(1.0)) <= 0.1


@pytest.mark.parametrize("kernel,deriv", [("b5", 0), ("b5", (1, 0, 2)), ("b7", (0, 3, 1)), ("b13", 0), ("b13", (1, 0, 0)), ("a4", 0)])
def test_fast_3d_agrees_with_direct_on_random_queries(kernel, deriv):
    """beyond the reference's single probe point: 3-D, derivatives, 4 000 random queries; the table + cubic-Hermite
    scheme is good to ~1e-9 of the data scale on a smooth field (the reference's own bound is 1e-6)"""
    rng = np.random.default_rng(7)
    g = np.linspace(0.0, 2 * np.pi, 24)
    v = np.sin(g)[:, None, None] * np.cos(g)[None, :, None] * np.sin(g + 0.3)[None, None, :]
    itp = O.Interpolant((g, g, g), v, kernel=kernel, derivative=deriv, bc="detect")
    q = rng.uniform(0.0, 2 * np.pi, (3, 4000))
    fast = itp.eval(*q)
    direct = O.direct_eval(itp, *q)
    scale = np.abs(itp.coefs).max() * np.prod([1.0 / h ** d for h, d in zip(itp.h, itp.derivs)])
    assert np.abs(fast - direct).max() <= 1e-6 * scale


def test_kernel_value_reproduces_the_shipped_tables():
    """K^(d)(s) by Horner in Float64 against the exactly rounded table entries, within the rounding bound of Horner's
    rule, 2 deg u sum_k |c_k| |s|^k (the high pieces of :b13 cancel 7 digits: 6e-9 absolute near |s| = 9)"""
    for kernel, d in (("b5", 0), ("b5", 1), ("b5", 3), ("b13", 0), ("b13", 6), ("a4", 0), ("a7", 1)):
        t = O.table(kernel, d)
        e = O.KERNEL_EQS[kernel]
        r = np.arange(101)[:, None] / 100.0
        s = r - e + np.arange(2 * e)[None, :]
        kv = O.kernel_value(kernel, d, s)
        # s = +-eqs is the edge of the support: the table holds the polynomial's limit there only where it is nonzero
        m = np.abs(s) < e
        _, pieces = O.kernel_poly_coefs(kernel, d)
        piece = np.minimum(np.floor(np.abs(s)).astype(int), e - 1)
        bound = np.zeros_like(s)
        for p, coef in enumerate(pieces):
            cond = sum(abs(c) * np.abs(s) ** k for k, c in enumerate(coef[d:]))
            # + the rounding of the argument itself, s = r/100 - eqs + c (a few ulp of eqs), through |dK/ds|
            dcond = sum(k * abs(c) * np.abs(s) ** (k - 1) for k, c in enumerate(coef[d:]) if k >= 1)
            bound = np.where(piece == p, 2 * len(coef) * 2.2e-16 * cond + 4 * 2.2e-16 * e * dcond, bound)
        assert (np.abs(kv - t)[m] <= bound[m] + 1e-300).all()
