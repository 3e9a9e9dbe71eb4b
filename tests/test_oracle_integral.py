"""The antiderivative fast path of the oracle (derivative = -1 on every axis) against the reference's own assertions
(test/test_antiderivative.jl: convergence to the analytic antiderivative, zero at the anchor) and against an independent
restatement of the tail arrays (numpy cumsum).  No GPU needed."""
import numpy as np
import pytest

KERNELS = ["a3", "a4", "a5", "a7", "b5", "b7", "b9", "b11", "b13"]
# test_antiderivative.jl expected_order_antideriv_fast is per kernel; every kernel is at least second order
MIN_RATIO = 2.0 ** 1


def _interior(r):
    h = r[1] - r[0]
    return np.concatenate([r[2:-2] + 0.25 * h, r[2:-2] + 0.7 * h])


@pytest.mark.parametrize("kernel", KERNELS)
def test_1d_antiderivative_of_sin_converges_and_is_zero_at_the_anchor(oracle, kernel):
    errs = []
    for n in (24, 48, 96):
        r = np.linspace(0.0, 2 * np.pi, n)
        itp = oracle.convolution_interpolation(r, np.sin(r), kernel=kernel, derivative=-1, bc="poly")
        a = itp.anchor[0]
        assert a == r[0]
        pts = _interior(r)
        errs.append(np.abs(itp.eval(pts) - (-np.cos(pts) + np.cos(a))).max())
        assert abs(itp(a)) < 1e-10  # "1D antiderivative anchor is zero"
    assert errs[0] / errs[1] > MIN_RATIO and errs[1] / errs[2] > MIN_RATIO
    assert errs[2] < 1e-6


@pytest.mark.parametrize("kernel", ["a4", "b5", "b9"])
@pytest.mark.parametrize("subgrid", ["cubic", "quintic"])
def test_2d_and_3d_product_antiderivative(oracle, kernel, subgrid):
    errs = []
    for n in (16, 32, 64):
        x = np.linspace(0.0, 2 * np.pi, n)
        v = np.sin(x)[:, None] * np.cos(x)[None, :]
        itp = oracle.convolution_interpolation((x, x), v, kernel=kernel, derivative=-1, bc="poly", subgrid=subgrid)
        ax, ay = itp.anchor
        q = np.random.default_rng(3).uniform(0.4, 5.8, (2, 400))
        exact = (-np.cos(q[0]) + np.cos(ax)) * (np.sin(q[1]) - np.sin(ay))
        errs.append(np.abs(itp.eval(*q) - exact).max())
    assert errs[0] / errs[1] > MIN_RATIO and errs[1] / errs[2] > MIN_RATIO
    errs = []
    for n in (12, 24, 48):
        x = np.linspace(0.0, 2 * np.pi, n)
        v = np.sin(x)[:, None, None] * np.cos(x)[None, :, None] * np.sin(x)[None, None, :]
        itp = oracle.convolution_interpolation((x, x, x), v, kernel=kernel, derivative=-1, bc="poly", subgrid=subgrid)
        ax, ay, az = itp.anchor
        q = np.random.default_rng(4).uniform(0.5, 5.7, (3, 300))
        exact = (-np.cos(q[0]) + np.cos(ax)) * (np.sin(q[1]) - np.sin(ay)) * (-np.cos(q[2]) + np.cos(az))
        errs.append(np.abs(itp.eval(*q) - exact).max())
    assert errs[0] / errs[1] > MIN_RATIO and errs[1] / errs[2] > MIN_RATIO


def test_a_series_integral_directions_resolve_to_the_cubic_subgrid(oracle):
    x = np.linspace(0, 1, 20)
    a = oracle.convolution_interpolation(x, x ** 2, kernel="a4", derivative=-1, subgrid="quintic")
    b = oracle.convolution_interpolation(x, x ** 2, kernel="b5", derivative=-1, subgrid="quintic")
    assert a.subgrid == ("cubic",) and b.subgrid == ("quintic",)
    m = oracle.convolution_interpolation((x, x), np.outer(x, x), kernel="b5", derivative=(-1, 0), subgrid="quintic")
    assert m.subgrid == ("quintic", "linear") and m.anchor == (0.0, 0.0)  # the other axis resolves to :linear


@pytest.mark.parametrize("kernel", ["b5", "a4"])
def test_mixed_integral_interpolation_axes(oracle, kernel):
    """FastMixedIntegralOrder: x-antiderivative / y-derivative of sin x cos y and the two-integral-axes 3-D case converge"""
    rng = np.random.default_rng(8)
    for dv, exact in (((-1, 0), lambda q, a: (-np.cos(q[0]) + np.cos(a[0])) * np.cos(q[1])),
                      ((0, -1), lambda q, a: np.sin(q[0]) * (np.sin(q[1]) - np.sin(a[1]))),
                      ((-1, 1), lambda q, a: (-np.cos(q[0]) + np.cos(a[0])) * (-np.sin(q[1])))):
        errs = []
        for n in (16, 32, 64):
            x = np.linspace(0, 2 * np.pi, n)
            itp = oracle.convolution_interpolation((x, x), np.sin(x)[:, None] * np.cos(x)[None, :], kernel=kernel,
                                                   derivative=dv, bc="poly")
            q = rng.uniform(0.5, 5.7, (2, 300))
            errs.append(np.abs(itp.eval(*q) - exact(q, itp.anchor)).max())
            assert itp.right_tail_maxabs() == 0.0
        assert errs[0] / errs[1] > MIN_RATIO and errs[1] / errs[2] > MIN_RATIO
    errs = []
    for n in (12, 24, 48):
        x = np.linspace(0, 2 * np.pi, n)
        v = np.sin(x)[:, None, None] * np.cos(x)[None, :, None] * np.sin(x)[None, None, :]
        itp = oracle.convolution_interpolation((x, x, x), v, kernel=kernel, derivative=(0, -1, -1), bc="poly")
        q = rng.uniform(0.6, 5.6, (3, 200))
        a = itp.anchor
        errs.append(np.abs(itp.eval(*q) - np.sin(q[0]) * (np.sin(q[1]) - np.sin(a[1])) * (-np.cos(q[2]) + np.cos(a[2]))).max())
    assert errs[0] / errs[1] > MIN_RATIO and errs[1] / errs[2] > MIN_RATIO


@pytest.mark.parametrize("shape,kernel", [((300,), "b5"), ((23, 19), ("b5", "a4")), ((14, 12, 13), ("a4", "b5", "b7"))])
def test_tail_arrays_against_numpy_and_right_tails_vanish(oracle, shape, kernel):
    """_build_tails_dispatch (constructor.jl:196-364) restated with numpy: Base cumsum along a dimension of an N-D array is
    the sequential np.cumsum; a Vector takes the pairwise route (<= 1 ulp-level differences)."""
    rng = np.random.default_rng(11)
    nd = len(shape)
    k = tuple(np.linspace(0, 1, s) for s in shape)
    v = rng.random(shape) + 0.5
    itp = oracle.convolution_interpolation(k if nd > 1 else k[0], v, kernel=kernel, derivative=-1, bc="detect")
    assert itp.right_tail_maxabs() == 0.0
    c = itp.coefs
    lv = [itp.left_values(d) for d in range(nd)]
    for d in range(nd):
        e = itp.eqs[d]
        assert np.all(lv[d][2 * e - 1:] == -0.5) and abs(lv[d][e - 1]) == 0.0  # K~(0) = 0 at the anchor
    for code in range(1, 3 ** nd):
        side = [(code // 3 ** d) % 3 for d in range(nd)]
        w = c.copy()
        for d in range(nd):
            if side[d]:
                sh = [1] * nd
                sh[d] = -1
                w = w * ((0.5 if side[d] == 1 else -0.5) - lv[d]).reshape(sh)
        if 2 in side:
            continue  # scan order of the suffix arrays follows the reference's own (oracle); they vanish where read
        for d in range(nd):
            if side[d] == 1:
                w = np.cumsum(w, axis=d)
        got = itp.tail(code)
        if nd == 1:
            assert np.abs(got - w).max() <= 1e-14 * np.abs(w).max()  # sequential vs pairwise association
        else:
            assert np.array_equal(got, w)
