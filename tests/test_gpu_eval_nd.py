"""Parity of the N >= 4 point-evaluation path (cib200_itpnd_*: eager_nd.jl:97-140 / perdim_nd.jl:4-36 behind the
extrapolation wrapper) with the CPU oracle, same seeded inputs, same tolerances as tests/test_gpu_eval.py."""
import numpy as np
import pytest

from .conftest import rel_err
from .test_gpu_eval import TOL32, TOL64, _compare, _field, _line_amplification, _scale

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel,n", [("b5", 14), ("a4", 12), ("b13", 20), ("a3", 10), ("b9", 18)])
@pytest.mark.parametrize("extrap", ["throw", "flat", "line"])
def test_4d_one_kernel(ci, oracle, kernel, n, extrap):
    rng = np.random.default_rng(401)
    shape = (n, n + 1, n + 2, n)
    k = tuple(np.linspace(0, 2 * np.pi, s) for s in shape)
    v = _field(shape, rng)
    lo, hi = (0.0, 2 * np.pi) if extrap == "throw" else (-0.3, 2 * np.pi + 0.3)
    q = rng.uniform(lo, hi, (4, 3000))
    _, _, g, _ = _compare(ci, oracle, k, v, tuple(q), kernel=kernel, extrap=extrap, bc="detect")
    assert g.itp.subgrid == ("linear",) * 4 and not g.itp.perdim


def test_4d_per_axis_kernels_and_derivatives(ci, oracle):
    rng = np.random.default_rng(402)
    shape = (16, 14, 15, 13)
    k = tuple(np.linspace(0, 1 + 0.1 * i, s) for i, s in enumerate(shape))
    v = _field(shape, rng)
    q = np.stack([rng.uniform(-0.05, kk[-1] + 0.05, 4000) for kk in k])
    for kw in (dict(kernel=("b5", "a4", "b7", "b5")), dict(kernel="b5", derivative=(1, 0, 0, 2)),
               dict(kernel=("b7", "b5", "b5", "b13"), derivative=(0, 3, 1, 0)), dict(kernel="b7", derivative=2),
               dict(kernel="b5", derivative=3)):
        for extrap in ("flat", "line"):
            _, _, g, _ = _compare(ci, oracle, k, v, tuple(q), extrap=extrap, bc="poly", **kw)
    assert g.itp.subgrid == ("linear",) * 4


def test_5d_and_6d_small(ci, oracle):
    rng = np.random.default_rng(403)
    shape = (9, 8, 10, 8, 9)
    k = tuple(np.linspace(0, 1, s) for s in shape)
    v = rng.random(shape)
    q = rng.uniform(-0.02, 1.02, (5, 1500))
    _compare(ci, oracle, k, v, tuple(q), kernel="a4", extrap="line")
    _compare(ci, oracle, k, v, tuple(q), kernel=("a4", "b5", "a3", "a4", "a5"), extrap="flat", bc="linear")
    shape = (7, 7, 6, 7, 6, 7)
    k = tuple(np.linspace(0, 1, s) for s in shape)
    v = rng.random(shape)
    q = rng.uniform(0, 1, (6, 500))
    _compare(ci, oracle, k, v, tuple(q), kernel="a3")


def test_4d_edge_queries_errors_and_cartesian_call(ci, oracle):
    rng = np.random.default_rng(404)
    x = np.linspace(0, 1, 12)
    k = (x, x, x, x)
    v = rng.random((12,) * 4)
    edge = np.concatenate([x, np.nextafter(x, 2), np.nextafter(x, -2)])
    edge = edge[(edge >= 0) & (edge <= 1)]
    q = np.stack([rng.choice(edge, 2000) for _ in range(4)])
    _compare(ci, oracle, k, v, tuple(q), kernel="b5")
    g = ci.convolution_interpolation(k, v, kernel="a4")
    qq = np.full((4, 5000), 0.5)
    qq[2, 4321] = 1.5
    with pytest.raises(ci.DomainError, match="query 4321 "):
        g.eval(*qq)
    qq[0, 17] = np.nan
    with pytest.raises(ci.InexactError, match="query 17 "):
        g.eval(*qq)
    assert g.eval(*[np.empty(0)] * 4).shape == (0,)
    gl = ci.convolution_interpolation(k, v, kernel="a4", extrap="line")
    o = oracle.convolution_interpolation(k, v, kernel="a4", extrap="line")
    xs = (rng.uniform(-0.1, 1.1, 5), rng.uniform(0, 1, 4), rng.uniform(0, 1, 3), rng.uniform(-0.1, 1.1, 6))
    got, ref = gl(*xs), o(*xs)
    assert got.shape == (5, 4, 3, 6)
    amp = _line_amplification(o, [a.ravel() for a in np.meshgrid(*xs, indexing="ij")]).reshape(got.shape)
    assert rel_err(got, ref, amp).max() <= TOL64
    assert gl(0.3, 0.9, 0.5, 0.1) == pytest.approx(o(0.3, 0.9, 0.5, 0.1), rel=1e-12)


def test_4d_device_resident_and_float32(ci, oracle):
    import torch
    rng = np.random.default_rng(405)
    shape = (12, 12, 12, 12)
    k = tuple(np.linspace(0, 1, s) for s in shape)
    v = _field(shape, rng)
    g = ci.convolution_interpolation(k, v, kernel="b5", extrap="flat")
    q = rng.uniform(-0.05, 1.05, (4, 20_000))
    out = g.eval(*[torch.from_numpy(a).cuda() for a in q])
    host = g.eval(*q)
    assert out.is_cuda and np.array_equal(out.cpu().numpy(), host)
    o = oracle.convolution_interpolation(k, v, kernel="b5", extrap="flat")
    ref = o.eval(*q)
    assert rel_err(host, ref, _scale(o)).max() <= TOL64
    g32 = ci.convolution_interpolation(k, v, kernel="b5", extrap="flat", dtype=np.float32)
    got32 = g32.eval(*q.astype(np.float32))
    assert got32.dtype == np.float32
    # Float32 queries are rounded before evaluation: compare against the oracle at the rounded coordinates
    ref32 = o.eval(*q.astype(np.float32).astype(np.float64))
    assert rel_err(got32, ref32, _scale(o)).max() <= TOL32


def test_4d_reference_extrapolation_property(ci):
    """test_extrapolation.jl:124-140 (4D fast kernels): a linear field is reproduced by :line and clamped by :flat"""
    N = 12
    r = np.linspace(0.0, 1.0, N)
    g = np.meshgrid(r, r, r, r, indexing="ij")
    vals = g[0] + g[1] + g[2] + g[3]
    pts = np.array([-0.2, -0.1, 1.1, 1.2])
    for kernel in ("a3", "a4", "b5", "b7"):
        line = ci.convolution_interpolation((r, r, r, r), vals, kernel=kernel, extrap="line")
        flat = ci.convolution_interpolation((r, r, r, r), vals, kernel=kernel, extrap="flat")
        assert np.allclose(line.eval(pts, pts, pts, pts), [-0.8, -0.4, 4.4, 4.8], atol=1e-6)
        assert np.allclose(flat.eval(pts, pts, pts, pts), [0.0, 0.0, 4.0, 4.0], atol=1e-6)
