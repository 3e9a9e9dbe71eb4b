"""Antiderivative / mixed interpolants in N >= 4 dimensions (the remainder):
convolution_fast_integration_nd.jl, convolution_fast_mixed_3_integral_dims.jl, convolution_fast_mixed_N_integral_dims.jl.
CPU: the numpy restatements against each other (the three-axis functor with all its tail arrays equals the brute-force
sum; the N-axis functor's truncated loops).  GPU: cib200_itgnd_* against the restatements."""
import numpy as np
import pytest

from .conftest import rel_err

ROWS = 1001  # table rows of the :linear axes in these tests (the reference's default 10 000 is used once, on the GPU)


def _case(shape, seed=3):
    rng = np.random.default_rng(seed)
    knots = tuple(np.linspace(0.0, 1.0 + 0.1 * d, n) for d, n in enumerate(shape))
    return knots, rng.random(shape), rng


def test_generated_antiderivative_table_reproduces_the_shipped_one(oracle):
    for k in ("b5", "b7", "b13", "a4"):
        assert np.array_equal(oracle.generated_table_i1(k, 101), oracle.table_i1(k))


@pytest.mark.parametrize("deriv", [(-1, -1, 0, -1), (-1, 1, -1, -1), (0, -1, -1, -1)])
def test_three_axis_functor_with_tails_equals_the_brute_force_sum(oracle, deriv):
    """convolution_fast_mixed_3_integral_dims.jl (27-region decomposition on prefix / suffix sums) and
    convolution_fast_mixed_N_integral_dims.jl (loop over the whole padded grid) state the same sum"""
    knots, v, rng = _case((9, 8, 10, 7))
    q = np.stack([rng.uniform(k[0], k[-1], 8) for k in knots])
    a = oracle.IntegralInterpolantND(knots, v, kernel="b5", derivative=deriv, bc="poly", linear_rows=ROWS)
    b = oracle.IntegralInterpolantND(knots, v, kernel="b5", derivative=deriv, bc="poly", linear_rows=ROWS, functor="mixedN")
    assert a.functor == "mixed3"
    ra, rb = a.eval(*q), b.eval(*q)
    assert np.abs(ra - rb).max() <= 1e-13 * max(1.0, np.abs(rb).max())


def test_n_axis_functor_visits_interior_extents_only(oracle):
    """convolution_fast_integration_nd.jl:20 loops over 1:domain_size[d] while indexing the padded arrays: with the loops
    extended to size(coefs, d) it would be the brute-force functor; as written it differs (quirk restated literally)"""
    knots, v, rng = _case((9, 8, 10, 7))
    q = np.stack([rng.uniform(k[0], k[-1], 5) for k in knots])
    a = oracle.IntegralInterpolantND(knots, v, kernel="b5", derivative=-1, bc="poly", linear_rows=ROWS)
    b = oracle.IntegralInterpolantND(knots, v, kernel="b5", derivative=-1, bc="poly", linear_rows=ROWS, functor="mixedN")
    assert a.functor == "nd"
    assert np.abs(a.eval(*q) - b.eval(*q)).max() > 1e-3
    # 3-D: the brute-force sum reproduces the tail-based 3-D functor of the C oracle (cubic vs linear subgrid: 1e-6)
    k3, v3, rng = _case((12, 11, 13))
    q3 = np.stack([rng.uniform(k[0], k[-1], 6) for k in k3])
    c = oracle.convolution_interpolation(k3, v3, kernel="b5", derivative=-1, bc="poly").eval(*q3)
    n = oracle.IntegralInterpolantND(k3, v3, kernel="b5", derivative=-1, bc="poly", functor="mixedN").eval(*q3)
    assert np.abs(c - n).max() <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("shape,kernel,deriv,extrap,rows", [
    ((9, 8, 10, 7), "b5", -1, "throw", ROWS),                      # every axis integral: nd functor (truncated loops)
    ((9, 8, 10, 7), "b5", (-1, -1, 0, -1), "line", ROWS),          # three integral axes + interpolation
    ((8, 9, 7, 10), ("b5", "b7", "b5", "a4"), (-1, 2, -1, -1), "flat", ROWS),  # ... + a derivative axis, mixed kernels
    ((7, 6, 8, 7, 6), "b5", (-1, -1, -1, 0, -1), "throw", ROWS),   # four integral axes of five: brute-force functor
    ((8, 7, 9, 8), "a4", -1, "throw", 10_000),                     # the reference's default table length
])
def test_cuda_nd_antiderivatives_against_the_restatement(ci, oracle, shape, kernel, deriv, extrap, rows):
    knots, v, rng = _case(shape, seed=11)
    lo, hi = (0.0, 1.0) if extrap == "throw" else (-0.08, 1.08)
    q = np.stack([k[0] + (k[-1] - k[0]) * rng.uniform(lo, hi, 40) for k in knots])
    o = oracle.IntegralInterpolantND(knots, v, kernel=kernel, derivative=deriv, bc="poly", extrap=extrap, linear_rows=rows)
    g = ci.convolution_interpolation(knots, v, kernel=kernel, derivative=deriv, bc="poly", extrap=extrap, linear_rows=rows)
    ref, got = o.eval(*q), g.eval(*q)
    # magnitude of the terms: max|coefs| x the volume the antiderivative covers x derivative scales
    S = np.abs(o.coefs).max() * abs(o._scale()) * np.prod([n if dv == -1 else 1 for n, dv in zip(shape, o.derivs)])
    amp = 1.0
    if extrap == "line":
        lo_b = np.array([k[0] for k in knots])[:, None]
        hi_b = np.array([k[-1] for k in knots])[:, None]
        dist = np.sqrt(((q - np.clip(q, lo_b, hi_b)) ** 2).sum(axis=0))
        amp = 1.0 + 2.0 * dist / (min(o.h) / 100.0)
    assert rel_err(got, ref, S * amp).max() <= 1e-12
    # Cartesian-product call
    gx = [np.sort(a[:3]) for a in q] if extrap != "throw" else [k[0] + (k[-1] - k[0]) * np.array([0.1, 0.5, 0.9]) for k in knots]
    grid = g(*gx)
    mesh = np.meshgrid(*gx, indexing="ij")
    flat = o.eval(*[m.ravel() for m in mesh]).reshape(grid.shape)
    amp2 = 1.0
    if extrap == "line":
        X = np.stack([m.ravel() for m in mesh])
        dist = np.sqrt(((X - np.clip(X, lo_b, hi_b)) ** 2).sum(axis=0)).reshape(grid.shape)
        amp2 = 1.0 + 2.0 * dist / (min(o.h) / 100.0)
    assert rel_err(grid, flat, S * amp2).max() <= 1e-12


@pytest.mark.gpu
def test_cuda_nd_antiderivative_modes_outside_the_path_are_rejected(ci):
    knots, v, _ = _case((8, 7, 9, 8))
    with pytest.raises(ci.UnsupportedModeError):
        ci.convolution_interpolation(knots, v, derivative=(-1, 0, 0, 0))       # one integral axis in 4-D
    with pytest.raises(ci.UnsupportedModeError):
        ci.convolution_interpolation(knots, v, derivative=(-1, -1, 0, 0))      # two
    with pytest.raises(Exception):
        ci.convolution_interpolation(knots, v, derivative=-1, extrap="throw").eval(*[np.array([2.0])] * 4)
