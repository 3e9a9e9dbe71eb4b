"""The reference's own analytic test assertions, run against the CUDA library, plus size-independent
invariants at sizes the oracle cannot reach (BASELINE config 5 geometry)."""
import types

import numpy as np
import pytest

from . import reference_properties as RP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def impl(ci):
    return types.SimpleNamespace(interp=ci.convolution_interpolation, resample=ci.convolution_resample,
                                 smooth=ci.convolution_smooth)


@pytest.mark.parametrize("check", RP.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_property(check, impl):
    check(impl)


def test_bc_downgrade_and_subgrid_resolution(ci):
    xs = np.linspace(0, 2 * np.pi, 5)
    itp = ci.convolution_interpolation(xs, np.sin(xs), kernel="b5", bc="poly")
    assert itp.itp.bc == (("linear", "linear"),)  # test_boundary_condition.jl:3-31
    v3 = np.sin(xs)[:, None, None] * np.cos(xs)[None, :, None] * np.sin(xs)[None, None, :]
    assert ci.convolution_interpolation((xs,) * 3, v3, kernel="b5", bc="poly").itp.bc == (("linear", "linear"),) * 3


@pytest.fixture(scope="module")
def big(ci):
    """256^3 :b5 grid (half of config 5's linear size; same code path), bc=:poly so that the map values -> result
    is linear, and 4*10^6 random queries."""
    rng = np.random.default_rng(0xC1B200 + 55)
    n = 256
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    g = np.linspace(0, 1, n)
    v1 = (np.sin(5 * g)[:, None, None] * np.cos(3 * g)[None, :, None] * np.sin(4 * g + 1)[None, None, :]
          + 0.01 * rng.uniform(-1, 1, (n, n, n)))
    v2 = rng.uniform(-1, 1, (n, n, n))
    q = rng.uniform(0, 1, (3, 4_000_000))
    return k, v1, v2, q


def test_large_grid_invariants(ci, big):
    k, v1, v2, q = big
    i1 = ci.convolution_interpolation(k, v1, bc="poly")
    r1 = i1.eval(*q)
    # (1) grid-point reproduction on a random subset of knots
    rng = np.random.default_rng(1)
    idx = rng.integers(0, 256, (3, 100_000))
    at = i1.eval(k[0][idx[0]], k[1][idx[1]], k[2][idx[2]])
    assert np.abs(at - v1[idx[0], idx[1], idx[2]]).max() < 1e-13
    # (2) permutation invariance, bitwise: results do not depend on where a query sits in the batch
    perm = rng.permutation(q.shape[1])
    rp = i1.eval(*q[:, perm])
    assert np.array_equal(rp, r1[perm])
    # (3) chunk invariance, bitwise
    half = q.shape[1] // 2 + 7
    assert np.array_equal(np.concatenate([i1.eval(*q[:, :half]), i1.eval(*q[:, half:])]), r1)
    # (4) linearity in the data (bc=:poly is a linear ghost rule)
    i2 = ci.convolution_interpolation(k, v2, bc="poly")
    i12 = ci.convolution_interpolation(k, 2.0 * v1 - 0.5 * v2, bc="poly")
    lin = 2.0 * r1 - 0.5 * i2.eval(*q)
    assert np.abs(i12.eval(*q) - lin).max() < 1e-12 * 4
    # (5) smooth data is interpolated to 7th order: error against the analytic field is tiny at n=256
    g = lambda a, b, c: np.sin(5 * a) * np.cos(3 * b) * np.sin(4 * c + 1)
    i0 = ci.convolution_interpolation(k, g(k[0][:, None, None], k[1][None, :, None], k[2][None, None, :]), bc="poly")
    assert np.abs(i0.eval(*q[:, :500_000]) - g(*q[:, :500_000])).max() < 1e-9
    for it in (i1, i2, i12, i0):
        it.itp.close()
