"""Parity of the antiderivative path (cib200_itg_*: convolution_fast_integration_{1d,2d,3d}.jl behind the extrapolation
wrapper, tails of fast_convolution_interpolation_constructor.jl:176-364) with the CPU oracle on the same seeded inputs.

Tolerance: |got - ref| <= 1e-12 * max(|ref|, S) with S = max|coefs| * prod(domain lengths) - the magnitude of the sum the
antiderivative accumulates (results are anchored to zero at the first knot, so they cross zero by construction); under
Line extrapolation the slope (f_b - f_in) / eps amplifies rounding by dist / eps like everywhere else."""
import numpy as np
import pytest

from .conftest import rel_err
from .test_gpu_eval import TOL64, _field, _line_amplification

pytestmark = pytest.mark.gpu

ALL_KERNELS = ["a3", "a4", "a5", "a7", "b5", "b7", "b9", "b11", "b13"]


def _scale(o):
    s = np.abs(o.coefs).max()
    for k in o.knots_in:
        s *= k[-1] - k[0]
    return s


def _compare(ci, oracle, knots, vals, xs, **kw):
    o = oracle.convolution_interpolation(knots, vals, derivative=-1, **kw)
    g = ci.convolution_interpolation(knots, vals, derivative=-1, **kw)
    ref, got = o.eval(*xs), g.eval(*xs)
    amp = _line_amplification(o, xs) if kw.get("extrap") == "line" else 1.0
    err = rel_err(got, ref, _scale(o) * amp)
    assert np.isfinite(got).all()
    assert err.max() <= TOL64, (err.max(), int(err.argmax()))
    assert g.itp.subgrid == o.subgrid and g.itp.anchor == o.anchor
    return got, ref, g, o


@pytest.mark.parametrize("kernel", ALL_KERNELS)
@pytest.mark.parametrize("subgrid", ["cubic", "quintic"])
def test_1d_all_kernels(ci, oracle, kernel, subgrid):
    rng = np.random.default_rng(701)
    for n in (40, 1000, 5000):  # 1000 / 5000: the pairwise prefix sum recurses (Base accumulate_pairwise!, blocks of 128)
        x = np.linspace(-1.0, 3.0, n)
        v = np.sin(3 * x) + 0.3 + 0.05 * rng.uniform(-1, 1, n)
        for extrap, lo, hi in (("throw", -1.0, 3.0), ("flat", -1.2, 3.2), ("line", -1.2, 3.2)):
            q = rng.uniform(lo, hi, 20_001)
            _, _, g, o = _compare(ci, oracle, x, v, (q,), kernel=kernel, subgrid=subgrid, extrap=extrap, bc="detect")
        assert np.array_equal(g.itp.left_values(0), o.left_values(0))
        t, tr = g.itp.tail(1), o.tail(1)
        assert np.abs(t - tr).max() <= 4e-16 * np.abs(tr).max()  # same pairwise order; device Hermite blend differs by an ulp
        assert abs(g(o.anchor[0])) <= 1e-13 * _scale(o)


@pytest.mark.parametrize("kw", [dict(kernel="b5"), dict(kernel="a4"), dict(kernel=("b13", "a3")), dict(kernel="b7", subgrid="quintic"),
                                dict(kernel=("b5", "b9"), subgrid=("quintic", "cubic"), bc="poly")])
@pytest.mark.parametrize("extrap", ["throw", "flat", "line"])
def test_2d(ci, oracle, kw, extrap):
    rng = np.random.default_rng(702)
    shape = (61, 47)
    k = (np.linspace(0, 2 * np.pi, shape[0]), np.linspace(-1, 1, shape[1]))
    v = _field(shape, rng)
    pad = 0.0 if extrap == "throw" else 0.2
    q = np.stack([rng.uniform(kk[0] - pad, kk[-1] + pad, 50_003) for kk in k])
    _, _, g, o = _compare(ci, oracle, k, v, tuple(q), extrap=extrap, **kw)
    for mask, code in ((1, 1), (2, 3), (3, 4)):  # tail1_left[1], tail1_left[2], tail2_ll
        t, tr = g.itp.tail(mask), o.tail(code)
        assert np.abs(t - tr).max() <= 1e-15 * np.abs(tr).max()


@pytest.mark.parametrize("kw", [dict(kernel="b5"), dict(kernel="a4"), dict(kernel=("a4", "b5", "b7"), bc="poly"),
                                dict(kernel="b13", subgrid="quintic")])
@pytest.mark.parametrize("extrap", ["throw", "line"])
def test_3d(ci, oracle, kw, extrap):
    rng = np.random.default_rng(703)
    shape = (33, 29, 31)
    k = tuple(np.linspace(0, 1 + 0.2 * i, s) for i, s in enumerate(shape))
    v = _field(shape, rng)
    pad = 0.0 if extrap == "throw" else 0.05
    q = np.stack([rng.uniform(kk[0] - pad, kk[-1] + pad, 30_001) for kk in k])
    _, _, g, o = _compare(ci, oracle, k, v, tuple(q), extrap=extrap, **kw)
    # faces, edges, corner: bit mask of the saturated axes -> base-3 code of the oracle (digit 1 = left)
    for mask in range(1, 8):
        code = sum(3 ** d for d in range(3) if mask >> d & 1)
        t, tr = g.itp.tail(mask), o.tail(code)
        assert np.abs(t - tr).max() <= 1e-15 * np.abs(tr).max()


def test_reference_assertions_on_the_device(ci):
    """test/test_antiderivative.jl "3D antiderivative convergence fast" and "anchor is zero", on the CUDA path"""
    errs = []
    for n in (12, 24, 48):
        x = np.linspace(0.0, 2 * np.pi, n)
        v = np.sin(x)[:, None, None] * np.cos(x)[None, :, None] * np.sin(x)[None, None, :]
        itp = ci.convolution_interpolation((x, x, x), v, kernel="b5", derivative=-1, bc="poly")
        ax, ay, az = itp.itp.anchor
        q = np.random.default_rng(5).uniform(0.0, 2 * np.pi, (3, 2000))
        exact = (-np.cos(q[0]) + np.cos(ax)) * (np.sin(q[1]) - np.sin(ay)) * (-np.cos(q[2]) + np.cos(az))
        errs.append(np.abs(itp.eval(*q) - exact).max())
        assert abs(itp(ax, ay, az)) < 1e-10
    assert errs[0] / errs[1] > 2.0 ** 4 and errs[1] / errs[2] > 2.0 ** 4 and errs[2] < 1e-6


def test_cartesian_call_device_tensors_and_edges(ci, oracle):
    import torch
    rng = np.random.default_rng(704)
    x, y = np.linspace(0, 1, 30), np.linspace(0, 2, 41)
    v = rng.random((30, 41))
    g = ci.convolution_interpolation((x, y), v, kernel="b5", derivative=-1, extrap="flat")
    o = oracle.convolution_interpolation((x, y), v, kernel="b5", derivative=-1, extrap="flat")
    S = _scale(o)
    qx = np.concatenate([x, np.nextafter(x, 2), np.nextafter(x, -2), [-0.1, 1.1]])
    qy = np.concatenate([y, np.nextafter(y, 3), np.nextafter(y, -3), [-0.1, 2.1]])
    got, ref = g(qx, qy), o(qx, qy)
    assert got.shape == (qx.size, qy.size) and rel_err(got, ref, S).max() <= TOL64
    q = rng.uniform(0, 1, (2, 10_000)) * np.array([[1.0], [2.0]])
    dq = [torch.from_numpy(a).cuda() for a in q]
    dev = g.eval(*dq)
    assert dev.is_cuda and rel_err(dev.cpu().numpy(), o.eval(*q), S).max() <= TOL64
    assert g.eval(np.empty(0), np.empty(0)).size == 0


def test_errors_and_unsupported_modes(ci):
    x = np.linspace(0, 1, 25)
    g = ci.convolution_interpolation(x, x ** 2, kernel="b5", derivative=-1)
    with pytest.raises(ci.DomainError):
        g.eval(np.array([0.5, 1.5, 0.2]))
    with pytest.raises(ci.InexactError):
        g.eval(np.array([0.5, np.nan]))
    v2 = np.outer(x, x)
    for kw in (dict(derivative=(-1, 4)), dict(derivative=-1, dtype=np.float32), dict(derivative=(-1, 0), dtype=np.float32)):
        with pytest.raises(ci.UnsupportedModeError):
            ci.convolution_interpolation((x, x), v2, kernel="b5", **kw)
    # N >= 4: three or more integral axes are accelerated (cib200_itgnd_*, tests/test_integral_nd.py); one or two are not
    with pytest.raises(ci.UnsupportedModeError):
        ci.convolution_interpolation((x[:9],) * 4, np.zeros((9,) * 4), kernel="b5", derivative=(-1, 0, -1, 0))
    with pytest.raises(ValueError):
        ci.convolution_interpolation(x, x, kernel="b5", derivative=-1, lazy=True)
    # a-series integral directions resolve to the cubic subgrid (_build_subgrids, constructor.jl:468-469)
    assert ci.convolution_interpolation(x, x, kernel="a4", derivative=-1, subgrid="quintic").itp.subgrid == ("cubic",)


def test_large_3d_grid_prefix_sums_and_linearity(ci):
    """size-independent properties at a size the oracle does not visit: the antiderivative is linear in the values, and
    the antiderivative of a constant c over the box [anchor, x] is c * volume"""
    rng = np.random.default_rng(705)
    n = 160
    x = np.linspace(0, 1, n)
    v1, v2 = rng.random((n, n, n)), rng.random((n, n, n))
    q = rng.uniform(0, 1, (3, 200_000))
    a = ci.convolution_interpolation((x, x, x), v1, kernel="b5", derivative=-1, bc="poly").eval(*q)
    b = ci.convolution_interpolation((x, x, x), v2, kernel="b5", derivative=-1, bc="poly").eval(*q)
    ab = ci.convolution_interpolation((x, x, x), 2.0 * v1 - 3.0 * v2, kernel="b5", derivative=-1, bc="poly").eval(*q)
    assert np.abs(ab - (2.0 * a - 3.0 * b)).max() <= 1e-12
    c = ci.convolution_interpolation((x, x, x), np.full((n, n, n), 1.75), kernel="b5", derivative=-1, bc="poly").eval(*q)
    assert np.abs(c - 1.75 * q[0] * q[1] * q[2]).max() <= 1e-12


@pytest.mark.parametrize("derivative,kw", [((-1, 0), dict(kernel="b5")), ((0, -1), dict(kernel="a4")),
                                          ((-1, 1), dict(kernel=("b7", "b5"), subgrid="quintic")),
                                          ((2, -1), dict(kernel="b5", bc="poly"))])
@pytest.mark.parametrize("extrap", ["throw", "flat", "line"])
def test_mixed_2d(ci, oracle, derivative, kw, extrap):
    """FastMixedIntegralOrder (convolution_fast_mixed_1_integral_dim.jl): one integral axis, the other on the :linear
    subgrid with the generated 10 000-row table"""
    rng = np.random.default_rng(711)
    shape = (53, 44)
    k = (np.linspace(0, 2 * np.pi, shape[0]), np.linspace(-1, 1, shape[1]))
    v = _field(shape, rng)
    pad = 0.0 if extrap == "throw" else 0.15
    q = np.stack([rng.uniform(kk[0] - pad, kk[-1] + pad, 40_003) for kk in k])
    o = oracle.convolution_interpolation(k, v, derivative=derivative, extrap=extrap, **kw)
    g = ci.convolution_interpolation(k, v, derivative=derivative, extrap=extrap, **kw)
    ref, got = o.eval(*q), g.eval(*q)
    S = np.abs(o.coefs).max()
    for kk, h, dv in zip(o.knots_in, o.h, derivative):
        S *= (kk[-1] - kk[0]) if dv == -1 else (1.0 / h) ** dv
    amp = _line_amplification(o, tuple(q)) if extrap == "line" else 1.0
    assert rel_err(got, ref, S * amp).max() <= TOL64
    assert g.itp.subgrid == o.subgrid and g.itp.anchor == o.anchor and "linear" in g.itp.subgrid


@pytest.mark.parametrize("derivative,kw", [((-1, 0, 0), dict(kernel="b5")), ((0, -1, -1), dict(kernel="b5", bc="poly")),
                                          ((-1, 1, -1), dict(kernel=("a4", "b5", "b7"))), ((0, 0, -1), dict(kernel="a4"))])
def test_mixed_3d(ci, oracle, derivative, kw):
    """one (mixed_1_integral_dim.jl) and two (mixed_2_integral_dims.jl) integral axes in 3-D"""
    rng = np.random.default_rng(712)
    shape = (27, 31, 25)
    k = tuple(np.linspace(0, 1 + 0.3 * i, s) for i, s in enumerate(shape))
    v = _field(shape, rng)
    for extrap, pad in (("throw", 0.0), ("line", 0.04)):
        q = np.stack([rng.uniform(kk[0] - pad, kk[-1] + pad, 20_001) for kk in k])
        o = oracle.convolution_interpolation(k, v, derivative=derivative, extrap=extrap, **kw)
        g = ci.convolution_interpolation(k, v, derivative=derivative, extrap=extrap, **kw)
        S = np.abs(o.coefs).max()
        for kk, h, dv in zip(o.knots_in, o.h, derivative):
            S *= (kk[-1] - kk[0]) if dv == -1 else (1.0 / h) ** dv
        amp = _line_amplification(o, tuple(q)) if extrap == "line" else 1.0
        assert rel_err(g.eval(*q), o.eval(*q), S * amp).max() <= TOL64


def test_mixed_analytic(ci):
    """d/dy of the x-antiderivative of sin x cos y = (-cos x + cos a) (-sin y), on the device path"""
    errs = []
    for n in (16, 32, 64):
        x = np.linspace(0.0, 2 * np.pi, n)
        itp = ci.convolution_interpolation((x, x), np.sin(x)[:, None] * np.cos(x)[None, :], kernel="b5",
                                           derivative=(-1, 1), bc="poly")
        q = np.random.default_rng(6).uniform(0.5, 5.7, (2, 1000))
        errs.append(np.abs(itp.eval(*q) - (-np.cos(q[0]) + np.cos(itp.itp.anchor[0])) * (-np.sin(q[1]))).max())
    assert errs[0] / errs[1] > 8 and errs[1] / errs[2] > 8 and errs[2] < 1e-7


def test_1d_long_grid_pairwise_recursion(ci, oracle):
    """n = 200 001: Base's accumulate_pairwise! recurses 11 levels (blocks of < 128); the device restates it in one thread
    with a raised stack limit - bit-for-bit the same association as the oracle's restatement"""
    rng = np.random.default_rng(706)
    n = 200_001
    x = np.linspace(0.0, 50.0, n)
    v = np.sin(x) + 0.5 + 0.01 * rng.uniform(-1, 1, n)
    q = rng.uniform(0.0, 50.0, 100_000)
    _, _, g, o = _compare(ci, oracle, x, v, (q,), kernel="b5", bc="poly")
    t, tr = g.itp.tail(1), o.tail(1)
    assert np.abs(t - tr).max() <= 4e-16 * np.abs(tr).max()
    exact = -np.cos(q) + np.cos(0.0) + 0.5 * q
    assert np.abs(g.eval(q) - exact).max() < 5e-3  # the integrated noise is a random walk of ~1e-3
