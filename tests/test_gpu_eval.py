"""Parity of the CUDA point-evaluation path (through the C-ABI) with the CPU oracle on the same seeded inputs.

Tolerance (north_star): <= 1e-12 relative in Float64, <= 1e-5 in Float32.  "Relative" is measured against
max(|ref|, S) with S = max|coefs| * prod_d (1/h_d)^DO_d - the magnitude of the terms being summed - because
results of sin-product fixtures cross zero.  For Line-extrapolated queries the
reference formula f_b + (f_b - f_in)/eps * dist amplifies ANY rounding difference by dist/eps, so S is scaled
by (1 + 2 dist/eps) for those queries; the amplification is a property of the formula, not of either side.
"""
import os

import numpy as np
import pytest

from .conftest import rel_err

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5


def _field(shape, rng, noise=0.05):
    grids = np.meshgrid(*[np.linspace(0, 2 * np.pi, n) for n in shape], indexing="ij")
    f = np.ones(shape)
    for i, g in enumerate(grids):
        f = f * (np.sin(g + 0.3 * i) if i % 2 == 0 else np.cos(g))
    return f + noise * rng.uniform(-1, 1, shape)


def _scale(oitp):
    s = np.abs(oitp.coefs).max()
    for h, d in zip(oitp.h, oitp.derivs):
        s *= (1.0 / h) ** d
    return s


def _line_amplification(oitp, xs):
    lo = np.array([k[e - 1] for k, e in zip(oitp.knots, oitp.eqs)])
    hi = np.array([k[-e] for k, e in zip(oitp.knots, oitp.eqs)])
    X = np.stack(xs)
    cl = np.clip(X, lo[:, None], hi[:, None])
    dist = np.sqrt(((X - cl) ** 2).sum(axis=0))
    return 1.0 + 2.0 * dist / (min(oitp.h) / 100.0)


def _compare(ci, oracle, knots, vals, xs, tol=TOL64, dtype=None, **kw):
    o = oracle.convolution_interpolation(knots, vals, **kw)
    g = ci.convolution_interpolation(knots, vals, dtype=dtype, **kw)
    ref = o.eval(*xs)
    got = g.eval(*xs)
    S = _scale(o)
    amp = _line_amplification(o, xs) if kw.get("extrap") == "line" else 1.0
    err = rel_err(got, ref, S * amp)
    assert np.isfinite(got).all()
    assert err.max() <= tol, (err.max(), int(err.argmax()))
    return got, ref, g, o


def test_c1_1d_b5_n1000_1e6_queries(ci, oracle):
    """BASELINE config 1 at full size."""
    rng = np.random.default_rng(0xC1B200 + 1)
    x = np.linspace(0, 1, 1000)
    v = np.sin(2 * np.pi * x) + 0.1 * rng.uniform(-1, 1, 1000)
    q = rng.uniform(0, 1, 1_000_000)
    for bc in ("detect", "poly"):
        _compare(ci, oracle, x, v, (q,), kernel="b5", bc=bc)


def test_c2_2d_b5_derivative_line_extrapolation(ci, oracle):
    """BASELINE config 2 at FULL size (200^2, derivative=(1,0) -> per-dim path, Line extrapolation, all 10^7 queries
    against the oracle)."""
    rng = np.random.default_rng(0xC1B200 + 2)
    x = np.linspace(0, 2 * np.pi, 200)
    v = np.sin(x)[:, None] * np.cos(x)[None, :] + 0.05 * rng.uniform(-1, 1, (200, 200))
    oracle.set_threads(os.cpu_count() or 1)
    q = rng.uniform(-0.1 * 2 * np.pi, 1.1 * 2 * np.pi, (2, 10_000_000))
    got, ref, g, o = _compare(ci, oracle, (x, x), v, (q[0], q[1]), kernel="b5", derivative=(1, 0), extrap="line")
    assert o.perdim and g.itp.perdim
    outside = (q < 0).any(axis=0) | (q > 2 * np.pi).any(axis=0)
    assert 0.2 < outside.mean() < 0.4  # ~31 % of the queries exercise the Line branch
    _compare(ci, oracle, (x, x), v, (q[0][:200_000], q[1][:200_000]), kernel="b5", derivative=(1, 0),
             extrap="line", bc="linear")


@pytest.mark.parametrize("extrap", ["throw", "flat", "line"])
def test_c5_small_3d_b5(ci, oracle, extrap):
    """BASELINE config 5 at oracle-sized dimensions: 3-D :b5, 64^3 grid, :detect ghosts, warp-cooperative kernel."""
    rng = np.random.default_rng(0xC1B200 + 5)
    shape = (64, 60, 56)
    knots = tuple(np.linspace(0, 1, n) for n in shape)
    v = _field(shape, rng, noise=0.01)
    lo, hi = (0.0, 1.0) if extrap == "throw" else (-0.08, 1.08)
    q = rng.uniform(lo, hi, (3, 100_003))  # ragged: not a multiple of the 32-query warp tile
    _compare(ci, oracle, knots, v, tuple(q), kernel="b5", extrap=extrap)


def test_c5_full_grid_512_cube_2e26_queries_against_the_oracle(ci, oracle):
    """BASELINE config 5 on its REAL grid: 512^3 :b5 through AUTO, device-resident, three densities: 4.2 10^8 + 9 uniformly
    random queries (3.1 per cell: tile binning + cell sort with group list + the tensor-core kernel on cell pairs; 32 768
    tiles, u32 record offsets, ragged groups), 2^26 (0.5 per cell: the tensor-core kernel on 2 x 2 x 2 cell blocks) and
    2^22 + 5 (binned kernel without the cell sort); a strided subsample of >= 2^20 results against the oracle on the same
    grid, and every result finite."""
    import torch
    rng = np.random.default_rng(0xC1B200 + 5)
    n = 512
    g = np.linspace(0.0, 1.0, n)
    a = np.sin(2 * np.pi * g) + 0.5 * np.cos(6 * np.pi * g + 0.3)
    b = np.cos(2 * np.pi * g) + 0.3 * np.sin(10 * np.pi * g)
    c = np.sin(4 * np.pi * g + 1.0)
    v = np.empty((n, n, n), order="F")
    for k in range(n):
        v[:, :, k] = a[:, None] * b[None, :] * c[k] + 0.01 * rng.uniform(-1, 1, (n, n))
    oracle.set_threads(os.cpu_count() or 1)
    o = oracle.convolution_interpolation((g, g, g), v, kernel="b5", bc="poly")
    itp = ci.convolution_interpolation((g, g, g), v, kernel="b5", bc="poly")
    assert np.array_equal(itp.itp.coefs, o.coefs) or np.abs(itp.itp.coefs - o.coefs).max() <= 1e-13 * np.abs(o.coefs).max()
    S = _scale(o)
    for nq, stride, strategy in ((420_000_009, 256, "cellmma"), (1 << 26, 64, "cellmma_block"), ((1 << 22) + 5, 4, "binned")):
        gen = torch.Generator(device="cuda")
        gen.manual_seed(nq)
        xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
        out = itp.eval(*xs)
        assert itp.itp.last_strategy() == strategy
        assert bool(torch.isfinite(out).all())
        sub = slice(0, nq, stride)
        q = [x[sub].cpu().numpy() for x in xs]
        ref = o.eval(*q)
        err = rel_err(out[sub].cpu().numpy(), ref, S)
        assert err.max() <= TOL64, (nq, err.max(), int(err.argmax()))
        assert len(ref) >= 1 << 20
        del xs, out


def test_auto_mode_falls_back_when_the_tables_do_not_fit_next_to_the_brick(ci, oracle):
    """:b13 with three different derivative orders stages six distinct 14.5 kB tables: brick + tables + scratch exceed
    the shared memory of an SM, so a batch above the AUTO threshold (2^22 queries) must take the gather kernel
    instead of failing at launch time (the same interpolant works for small batches)."""
    import torch
    rng = np.random.default_rng(5)
    g = np.linspace(0.0, 1.0, 28)
    v = _field((28, 28, 28), rng, noise=0.01)
    kw = dict(kernel="b13", derivative=(0, 2, 4), bc="detect")
    o = oracle.convolution_interpolation((g, g, g), v, **kw)
    itp = ci.convolution_interpolation((g, g, g), v, **kw)
    nq = (1 << 22) + 3
    gen = torch.Generator(device="cuda")
    gen.manual_seed(9)
    xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
    out = itp.eval(*xs)
    sub = slice(0, nq, 97)
    ref = o.eval(*[x[sub].cpu().numpy() for x in xs])
    err = rel_err(out[sub].cpu().numpy(), ref, _scale(o))
    assert err.max() <= TOL64, err.max()


@pytest.mark.parametrize("kernel,deriv", [("b5", 0), ("b5", (1, 0, 2)), ("b7", (0, 3, 1)), ("b13", 0), ("b13", (1, 0, 0)), ("a4", 0)])
def test_cuda_path_agrees_with_the_direct_evaluator(ci, oracle, kernel, deriv):
    """Second, independent oracle: the reference's DIRECT evaluators (kernel polynomials by Horner, no tables;
    oracle.direct_eval, convolution_kernel_interpolation_eager_3d.jl:1-29).  The reference asserts fast ~ direct to 1e-6
    (test/test_constructors.jl:21,45); the CUDA library is held to the same bound against the direct statement."""
    rng = np.random.default_rng(7)
    g = np.linspace(0.0, 2 * np.pi, 24)
    v = np.sin(g)[:, None, None] * np.cos(g)[None, :, None] * np.sin(g + 0.3)[None, None, :]
    q = rng.uniform(0.0, 2 * np.pi, (3, 4000))
    o = oracle.Interpolant((g, g, g), v, kernel=kernel, derivative=deriv, bc="detect")
    got = ci.convolution_interpolation((g, g, g), v, kernel=kernel, derivative=deriv, bc="detect").eval(*q)
    direct = oracle.direct_eval(o, *q)
    assert np.abs(got - direct).max() <= 1e-6 * _scale(o)
    # 1-D and 2-D, every kernel of the reference's test loop (test_constructors.jl:9)
    x = np.linspace(0.0, 2 * np.pi, 30)
    q1 = rng.uniform(0.0, 2 * np.pi, 3000)
    o1 = oracle.Interpolant((x,), np.sin(x), kernel=kernel)
    g1 = ci.convolution_interpolation(x, np.sin(x), kernel=kernel).eval(q1)
    assert np.abs(g1 - oracle.direct_eval(o1, q1)).max() <= 1e-6
    v2 = np.sin(x)[:, None] * np.cos(x)[None, :]
    o2 = oracle.Interpolant((x, x), v2, kernel=kernel)
    q2 = rng.uniform(0.0, 2 * np.pi, (2, 3000))
    g2 = ci.convolution_interpolation((x, x), v2, kernel=kernel).eval(*q2)
    assert np.abs(g2 - oracle.direct_eval(o2, *q2)).max() <= 1e-6


@pytest.mark.parametrize("kernel,deriv,bc", [("b13", (1, 0, 0), "detect"), ("b13", 0, "poly"), ("b7", (0, 2, 1), "linear"),
                                             ("b9", 1, "quadratic"), ("b11", (0, 0, 3), "detect"), ("b5", (2, 2, 2), "poly")])
def test_3d_kernels_and_derivatives(ci, oracle, kernel, deriv, bc):
    rng = np.random.default_rng(11)
    shape = (26, 24, 28)
    knots = (np.linspace(0, 1, 26), np.linspace(-1, 1, 24), np.linspace(2, 5, 28))
    v = _field(shape, rng)
    q = np.stack([rng.uniform(k[0], k[-1], 8000) for k in knots])
    _compare(ci, oracle, knots, v, tuple(q), kernel=kernel, derivative=deriv, bc=bc)


def test_mixed_kernels_per_axis_use_the_generic_path(ci, oracle):
    rng = np.random.default_rng(12)
    shape = (22, 25, 21)
    knots = tuple(np.linspace(0, 1, n) for n in shape)
    v = _field(shape, rng)
    q = rng.uniform(0, 1, (3, 5000))
    _compare(ci, oracle, knots, v, tuple(q), kernel=("b5", "b13", "b7"), derivative=(0, 1, 0), extrap="line")
    v2 = _field((30, 33), rng)
    k2 = (np.linspace(0, 1, 30), np.linspace(0, 2, 33))
    q2 = np.stack([rng.uniform(-0.1, 1.1, 20000), rng.uniform(-0.1, 2.1, 20000)])
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=("b13", "b5"), extrap="flat")
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=("b7", "b7"), derivative=(1, 1), extrap="line")


@pytest.mark.parametrize("kernel", ["a3", "a4", "a5", "a7"])
@pytest.mark.parametrize("bc", ["detect", "poly", "linear", "quadratic"])
def test_a_series_kernels_1d_2d_3d(ci, oracle, kernel, bc):
    """a-series kernels (a4 is the reference's default): derivative 0, cubic subgrid, 4-node polynomial ghosts."""
    rng = np.random.default_rng(41)
    x = np.linspace(-2, 3, 61)
    v = np.cos(1.7 * x) + 0.03 * rng.standard_normal(61)
    q = rng.uniform(-2.3, 3.3, 40_000)
    _compare(ci, oracle, x, v, (q,), kernel=kernel, bc=bc, extrap="line")
    v2 = _field((40, 35), rng, noise=0.02)
    k2 = (np.linspace(0, 1, 40), np.linspace(0, 1, 35))
    q2 = rng.uniform(-0.05, 1.05, (2, 30_000))
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel, bc=bc, extrap="flat")
    shape = (26, 24, 28)
    k3 = tuple(np.linspace(0, 1, n) for n in shape)
    v3 = _field(shape, rng, noise=0.02)
    q3 = rng.uniform(0, 1, (3, 20_011))
    _compare(ci, oracle, k3, v3, tuple(q3), kernel=kernel, bc=bc)
    if bc == "detect":
        # mixed a / b kernels per axis take the generic per-axis path; derivatives of a-kernels are not shipped
        _compare(ci, oracle, k3, v3, tuple(q3), kernel=(kernel, "b5", "a4"), derivative=(0, 1, 0), extrap="line")
        # derivative 1 of an a-kernel resolves to the :linear subgrid (K' rows, linear blend)
        g = _compare(ci, oracle, x, v, (q,), kernel=kernel, derivative=1, extrap="line")[2]
        assert g.itp.subgrid == ("linear",)
        with pytest.raises(ci.UnsupportedModeError):
            ci.convolution_interpolation(x, v, kernel=kernel, derivative=2)


@pytest.mark.parametrize("kernel", ["b5", "b9", "b13"])
def test_linear_subgrid_and_top_derivatives(ci, oracle, kernel):
    """:linear subgrid (K rows only, linear blend; precompute = 101): requested explicitly, reached through a top
    derivative (b5: 3, b9: 5, b13: 6), mixed with cubic / quintic axes on the per-dim 2-D path, the all-linear 3-D
    evaluators (any non-all-cubic 3-D interpolant), and the (linear, linear) position arithmetic of eager_2d.jl."""
    rng = np.random.default_rng(43)
    top = {"b5": 3, "b9": 5, "b13": 6}[kernel]
    x = np.linspace(-2, 3, 77)
    v = np.cos(1.7 * x) + 0.03 * rng.standard_normal(77)
    q = rng.uniform(-2.3, 3.3, 40_000)
    _compare(ci, oracle, x, v, (q,), kernel=kernel, subgrid="linear", extrap="line")
    g = _compare(ci, oracle, x, v, (q,), kernel=kernel, derivative=top, extrap="flat")[2]
    assert g.itp.subgrid == ("linear",)
    v2 = _field((40, 35), rng, noise=0.02)
    k2 = (np.linspace(0, 1, 40), np.linspace(0, 1, 35))
    q2 = rng.uniform(-0.05, 1.05, (2, 30_000))
    g = _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel, subgrid="linear", extrap="line")[2]   # eager (linear, linear)
    assert g.itp.subgrid == ("linear", "linear") and not g.itp.perdim
    g = _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel, derivative=(top, 0), extrap="flat")[2]  # per-dim (linear, cubic)
    assert g.itp.subgrid == ("linear", "cubic") and g.itp.perdim
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=(kernel, "b7"), derivative=(0, 1), subgrid=("quintic", "linear"),
             extrap="flat")
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=(kernel, "b7"), subgrid=("linear", "linear"), extrap="line")
    shape = (26, 24, 28)
    k3 = tuple(np.linspace(0, 1, n) for n in shape)
    v3 = _field(shape, rng, noise=0.02)
    q3 = rng.uniform(-0.03, 1.03, (3, 20_011))
    g = _compare(ci, oracle, k3, v3, tuple(q3), kernel=kernel, derivative=(top, 0, 0), extrap="flat")[2]
    assert g.itp.subgrid == ("linear",) * 3 and g.itp.perdim
    _compare(ci, oracle, k3, v3, tuple(q3), kernel=kernel, subgrid="linear", extrap="line")
    _compare(ci, oracle, k3, v3, tuple(q3), kernel=kernel, subgrid=("cubic", "quintic", "cubic"), extrap="flat")
    # the reference's :linear axes read generated 10 000-row tables (constructor.jl:117-118): both sides did above; on the
    # 101-row shipped tables the same interpolant is ~1e-4 away (and both sides agree on that variant too)
    g101 = ci.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", extrap="line", linear_rows=101)
    o101 = oracle.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", extrap="line", linear_rows=101)
    r101 = o101.eval(*q3)
    assert rel_err(g101.eval(*q3), r101, _scale(o101) * _line_amplification(o101, tuple(q3))).max() <= TOL64
    r10k = oracle.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", extrap="line").eval(*q3)
    inside = ((q3 >= 0) & (q3 <= 1)).all(axis=0)
    assert 1e-9 < np.abs(r101 - r10k)[inside].max() / np.abs(r10k)[inside].max() < 1e-2
    assert g101.itp.linear_rows == 101 and g.itp.linear_rows == 10_000
    if kernel == "b5":  # the binned strategies on an all-linear 3-D interpolant
        qb = rng.uniform(0, 1, (3, 300_007))
        gg = ci.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", eval_mode="gather", linear_rows=101)
        gb = ci.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", eval_mode="binned", linear_rows=101)
        assert rel_err(gb.eval(*qb), gg.eval(*qb), np.abs(v3).max()).max() <= 1e-14
        # generated 10 000-row tables: read from global memory by the binned kernel (round 2)
        gb10k = ci.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", eval_mode="binned")
        gg10k = ci.convolution_interpolation(k3, v3, kernel=kernel, subgrid="linear", eval_mode="gather")
        assert rel_err(gb10k.eval(*qb), gg10k.eval(*qb), np.abs(v3).max()).max() <= 1e-14


@pytest.mark.parametrize("kernel", ["b5", "b7", "b13"])
def test_quintic_subgrid_1d_2d(ci, oracle, kernel):
    """:quintic subgrid (K, K', K'' rows, quintic_hermite): 1-D, 2-D (quintic, quintic) on the eager path, the mixed
    (quintic, cubic) / (cubic, quintic) pairs on the per-dim path, derivatives while K^(d+2) is shipped."""
    rng = np.random.default_rng(42)
    x = np.linspace(-2, 3, 77)
    v = np.cos(1.7 * x) + 0.03 * rng.standard_normal(77)
    q = rng.uniform(-2.3, 3.3, 40_000)
    for d in (0, 1):
        g = _compare(ci, oracle, x, v, (q,), kernel=kernel, derivative=d, subgrid="quintic", extrap="line")[2]
        assert g.itp.subgrid == ("quintic",)
    v2 = _field((40, 35), rng, noise=0.02)
    k2 = (np.linspace(0, 1, 40), np.linspace(0, 1, 35))
    q2 = rng.uniform(-0.05, 1.05, (2, 30_000))
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel, subgrid="quintic", extrap="flat")
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel, derivative=(1, 1), subgrid="quintic", extrap="line")
    for sg in (("quintic", "cubic"), ("cubic", "quintic"), ("quintic", "quintic")):
        _compare(ci, oracle, k2, v2, tuple(q2), kernel=(kernel, "b9"), derivative=(1, 0), subgrid=sg, extrap="line")
    # quintic and cubic subgrids agree to the accuracy of the table interpolation, not to rounding
    gq = ci.convolution_interpolation(x, v, kernel=kernel, subgrid="quintic").eval(np.clip(q, -2, 3))
    gc = ci.convolution_interpolation(x, v, kernel=kernel).eval(np.clip(q, -2, 3))
    assert 0 < np.abs(gq - gc).max() < 1e-8


@pytest.mark.parametrize("kernel", ["b5", "b7", "b9", "b11", "b13"])
def test_1d_2d_all_b_kernels(ci, oracle, kernel):
    rng = np.random.default_rng(13)
    x = np.linspace(-2, 3, 77)
    v = np.cos(1.7 * x) + 0.03 * rng.standard_normal(77)
    q = rng.uniform(-2.3, 3.3, 50_000)
    for d in (0, 1, 2):
        _compare(ci, oracle, x, v, (q,), kernel=kernel, derivative=d, extrap="line")
    v2 = _field((40, 35), rng)
    k2 = (np.linspace(0, 1, 40), np.linspace(0, 1, 35))
    q2 = rng.uniform(0, 1, (2, 30_000))
    _compare(ci, oracle, k2, v2, tuple(q2), kernel=kernel)


def test_boundary_and_edge_queries(ci, oracle):
    """queries exactly on knots, on the domain bounds, within 1 ulp of cell boundaries, and far outside"""
    rng = np.random.default_rng(14)
    x = np.linspace(0, 1, 33)
    v = rng.random(33)
    edge = np.concatenate([x, np.nextafter(x, 2), np.nextafter(x, -2), [0.0, 1.0, -5.0, 7.5, 1e6, -1e6]])
    for extrap in ("flat", "line"):
        _compare(ci, oracle, x, v, (edge,), extrap=extrap)
    inside = edge[(edge >= 0) & (edge <= 1)]
    _compare(ci, oracle, x, v, (inside,), extrap="throw")
    k3 = (x, x, x)
    v3 = _field((33, 33, 33), rng)
    e3 = np.stack([rng.choice(inside, 4000), rng.choice(inside, 4000), rng.choice(inside, 4000)])
    _compare(ci, oracle, k3, v3, tuple(e3))


def test_throw_reports_first_offending_query(ci, oracle):
    x = np.linspace(0, 1, 50)
    g = ci.convolution_interpolation(x, np.sin(x))
    q = np.full(100_000, 0.5)
    q[[70_001, 90_000]] = 1.5
    with pytest.raises(ci.DomainError, match="query 70001 "):
        g.eval(q)
    q[123] = np.nan
    with pytest.raises(ci.InexactError, match="query 123 "):
        g.eval(q)
    g3 = ci.convolution_interpolation((x, x, x), np.zeros((50, 50, 50)), extrap="flat")
    q3 = np.full((3, 1000), 0.5)
    q3[1, 777] = np.inf  # Flat clamps +Inf; Line would produce NaN -> InexactError in the reference
    assert np.isfinite(g3.eval(*q3)).all()
    g3l = ci.convolution_interpolation((x, x, x), np.zeros((50, 50, 50)), extrap="line")
    with pytest.raises(ci.InexactError, match="query 777 "):
        g3l.eval(*q3)
    with pytest.raises(oracle.OracleError, match="query 777"):
        oracle.convolution_interpolation((x, x, x), np.zeros((50, 50, 50)), extrap="line").eval(*q3)


def test_empty_and_tiny_batches(ci, oracle):
    x = np.linspace(0, 1, 20)
    v3 = np.random.default_rng(3).random((20, 20, 20))
    g = ci.convolution_interpolation((x, x, x), v3)
    assert g.eval(np.empty(0), np.empty(0), np.empty(0)).shape == (0,)
    o = oracle.convolution_interpolation((x, x, x), v3)
    for n in (1, 2, 31, 32, 33):
        q = np.random.default_rng(n).uniform(0, 1, (3, n))
        assert rel_err(g.eval(*q), o.eval(*q), 1.0).max() <= TOL64


def test_cartesian_product_call(ci, oracle):
    rng = np.random.default_rng(15)
    k = (np.linspace(0, 1, 20), np.linspace(0, 2, 24), np.linspace(-1, 0, 22))
    v = rng.random((20, 24, 22))
    g = ci.convolution_interpolation(k, v, extrap="line")
    o = oracle.convolution_interpolation(k, v, extrap="line")
    xs = (rng.uniform(-0.1, 1.1, 17), rng.uniform(0, 2, 9), rng.uniform(-1, 0, 13))
    got, ref = g(*xs), o(*xs)
    assert got.shape == (17, 9, 13)
    amp = _line_amplification(o, [a.ravel() for a in np.meshgrid(*xs, indexing="ij")]).reshape(got.shape)
    assert rel_err(got, ref, amp).max() <= TOL64
    assert g(0.3, 1.1, -0.5) == pytest.approx(o(0.3, 1.1, -0.5), rel=1e-12)


def test_device_resident_queries(ci, oracle):
    import torch
    rng = np.random.default_rng(16)
    k = tuple(np.linspace(0, 1, 40) for _ in range(3))
    v = _field((40, 40, 40), rng)
    g = ci.convolution_interpolation(k, v, extrap="flat")
    q = rng.uniform(-0.05, 1.05, (3, 50_000))
    tq = [torch.from_numpy(a).cuda() for a in q]
    out = g.eval(*tq)
    assert out.is_cuda and out.dtype == torch.float64
    host = g.eval(*q)
    assert np.array_equal(out.cpu().numpy(), host)  # same kernel, same bits, whichever side the buffers live
    o = oracle.convolution_interpolation(k, v, extrap="flat")
    assert rel_err(host, o.eval(*q), _scale(o)).max() <= TOL64


def test_float32_opt_in_mode(ci, oracle):
    """Float32 mode has no same-precision reference (the reference's shipped tables are Float64 only,
    SURVEY.md 4); its bound is <= 1e-5 against the Float64 result."""
    rng = np.random.default_rng(17)
    k = tuple(np.linspace(0, 1, 48) for _ in range(3))
    v = _field((48, 48, 48), rng).astype(np.float32)
    q = rng.uniform(0, 1, (3, 40_000)).astype(np.float32)
    o = oracle.convolution_interpolation(k, v.astype(np.float64), bc="poly")
    g = ci.convolution_interpolation(k, v, bc="poly")
    assert g.itp.dtype == np.float32
    got = g.eval(*q)
    assert got.dtype == np.float32
    ref = o.eval(*[a.astype(np.float64) for a in q])
    assert rel_err(got, ref, _scale(o)).max() <= TOL32
    # Float32 knots must pass the reference's uniformity test at rtol 1e-5 (nonuniform.jl:13-25): a
    # 300-point Float32 linspace does NOT (spacing jitter 3.6e-5) and is routed to the non-uniform path by the
    # reference too; dyadic knots are exactly uniform in Float32
    x = (np.arange(257) / 256).astype(np.float32)
    with pytest.raises(ci.UnsupportedModeError, match="non-uniform"):
        ci.convolution_interpolation(np.linspace(0, 1, 300).astype(np.float32), np.zeros(300, np.float32))
    v1 = np.sin(6 * x).astype(np.float32)
    q1 = rng.uniform(0, 1, 10_000).astype(np.float32)
    g1 = ci.convolution_interpolation(x, v1, derivative=1, bc="poly")
    o1 = oracle.convolution_interpolation(x.astype(np.float64), v1.astype(np.float64), derivative=1, bc="poly")
    assert rel_err(g1.eval(q1), o1.eval(q1.astype(np.float64)), _scale(o1)).max() <= TOL32


@pytest.mark.parametrize("kernel,nq,strategy", [("b5", 200_003, "cellmma_block"), ("b5", (1 << 22) + 77, "cellmma"), ("b5", 40_000_003, "cellmma"),
                                                ("b13", (1 << 22) + 5, "binned")])
def test_float32_mode_on_the_binned_strategies(ci, oracle, kernel, nq, strategy):
    """Float32 interpolants take the tile-binned strategies too (the brick is staged as floats by a FLOAT32 tensor map,
    records and arithmetic stay double): a batch above the AUTO threshold - for :b5 the tensor-core kernels (float brick
    converted when a B fragment is loaded; cell blocks from 0.4 query per cell, cell pairs from 3) - against the Float64
    oracle (1e-5) and against the same interpolant's gather kernel."""
    import torch
    rng = np.random.default_rng(23)
    shape = (56, 64, 48)  # padded extents multiples of 4: Float32 TMA strides are multiples of 16 bytes
    knots = tuple((np.arange(n) / 64).astype(np.float32) for n in shape)
    v = _field(shape, rng, noise=0.01).astype(np.float32)
    extrap = "flat" if kernel == "b13" else "throw"
    itp = ci.convolution_interpolation(knots, v, kernel=kernel, bc="poly", extrap=extrap)
    assert itp.itp.dtype == np.float32
    o = oracle.convolution_interpolation(tuple(k.astype(np.float64) for k in knots), v.astype(np.float64), kernel=kernel,
                                         bc="poly", extrap=extrap)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(nq)
    hi = [float(k[-1]) for k in knots]
    if extrap == "flat":  # 2 % of the axis outside on either side
        xs = [(torch.rand(nq, dtype=torch.float32, device="cuda", generator=gen) * 1.04 - 0.02) * h for h in hi]
    else:
        xs = [(torch.rand(nq, dtype=torch.float32, device="cuda", generator=gen) * h).clamp(0.0, h) for h in hi]
    out = itp.eval(*xs)
    assert itp.itp.last_strategy() == strategy
    assert out.dtype == torch.float32 and bool(torch.isfinite(out).all())
    sub = slice(0, nq, max(1, nq // 200_000))
    q = [x[sub].cpu().numpy() for x in xs]
    ref = o.eval(*[a.astype(np.float64) for a in q])
    assert rel_err(out[sub].cpu().numpy(), ref, _scale(o)).max() <= TOL32
    small = itp.eval(*[x[sub].contiguous() for x in xs])  # below the threshold: gather kernel
    assert float((small - out[sub]).abs().max()) <= 1e-6 * _scale(o)


@pytest.mark.parametrize("kernel,shape", [("b5", (64, 60, 56)), ("b13", (40, 44, 36)), ("b7", (50, 30, 34)),
                                          ("a3", (50, 44, 60)), ("a4", (64, 60, 56)), ("a7", (40, 52, 36))])
def test_binned_strategy_equals_gather_strategy(ci, oracle, kernel, shape):
    """The binned strategy (tile binning + TMA-staged bricks) uses the same positions and the same per-axis
    weights as the direct gather kernel; only the association of the (2E)^3-term sum differs (lanes own
    different taps), so the two strategies agree to a few ulp of sum|c w| and both meet the parity bound against
    the oracle.  Queries that need extrapolation go through the spill kernel."""
    rng = np.random.default_rng(31)
    knots = tuple(np.linspace(0, 1, n) for n in shape)
    v = _field(shape, rng, noise=0.02)
    q = rng.uniform(-0.03, 1.03, (3, 300_007))
    for extrap in ("flat", "line"):
        gg = ci.convolution_interpolation(knots, v, kernel=kernel, extrap=extrap, eval_mode="gather")
        gb = ci.convolution_interpolation(knots, v, kernel=kernel, extrap=extrap, eval_mode="binned")
        rg, rb = gg.eval(*q), gb.eval(*q)
        inside = ((q >= 0) & (q <= 1)).all(axis=0)
        assert inside.sum() > 200_000 and (~inside).sum() > 10_000
        assert rel_err(rb[inside], rg[inside], np.abs(v).max()).max() <= 1e-14
        o = oracle.convolution_interpolation(knots, v, kernel=kernel, extrap=extrap)
        amp = _line_amplification(o, q) if extrap == "line" else 1.0
        assert rel_err(rb, o.eval(*q), _scale(o) * amp).max() <= TOL64
    # Throw: error reporting also works through the spill bin
    gb = ci.convolution_interpolation(knots, v, kernel=kernel, eval_mode="binned")
    qi = np.clip(q, 0, 1)
    rg = ci.convolution_interpolation(knots, v, kernel=kernel, eval_mode="gather").eval(*qi)
    assert rel_err(gb.eval(*qi), rg, np.abs(v).max()).max() <= 1e-14
    qi[1, 12345] = 1.5
    with pytest.raises(ci.DomainError, match="query 12345 "):
        gb.eval(*qi)
    qi[2, 777] = np.nan
    with pytest.raises(ci.InexactError, match="query 777 "):
        gb.eval(*qi)


def test_binned_strategy_with_more_tiles_than_a_shared_memory_histogram_holds(ci, oracle):
    """:b13 tiles are 8^3 cells: a 300^3 grid has 54 872 of them, more than the 49 152 bins of the per-CTA shared-memory
    histograms - the binning passes then keep one histogram / one write cursor per tile in global memory"""
    import torch
    rng = np.random.default_rng(77)
    n = 300
    g = np.linspace(0.0, 1.0, n)
    a, b, c = np.sin(2 * np.pi * g), np.cos(3 * np.pi * g), np.sin(np.pi * g + 0.5)
    v = a[:, None, None] * b[None, :, None] * c[None, None, :] + 0.01 * rng.uniform(-1, 1, (n, n, n))
    itp = ci.convolution_interpolation((g, g, g), v, kernel="b13", bc="poly", eval_mode="binned")
    o = oracle.convolution_interpolation((g, g, g), v, kernel="b13", bc="poly")
    nq = 3_000_017
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
    out = itp.eval(*xs)
    assert bool(torch.isfinite(out).all())
    sub = slice(0, nq, 151)
    ref = o.eval(*[x[sub].cpu().numpy() for x in xs])
    assert rel_err(out[sub].cpu().numpy(), ref, _scale(o)).max() <= TOL64


@pytest.mark.parametrize("kw", [dict(kernel="b5", derivative=(3, 0, 1)), dict(kernel="b7", subgrid=("cubic", "linear", "cubic")),
                                dict(kernel="b13", derivative=(6, 0, 0)), dict(kernel="a4", derivative=(1, 0, 0))])
def test_linear_subgrid_interpolants_on_the_binned_strategy(ci, oracle, kw):
    """top derivatives / non-all-cubic 3-D interpolants resolve to the :linear subgrid on generated 10 000-row tables
    (constructor.jl:117-118): the binned kernel reads those tables from global memory; same results as the gather
    kernel and as the oracle"""
    rng = np.random.default_rng(41)
    shape = (44, 40, 36)
    knots = tuple(np.linspace(0, 1, n) for n in shape)
    v = _field(shape, rng, noise=0.01)
    q = rng.uniform(-0.02, 1.02, (3, 200_003))
    gb = ci.convolution_interpolation(knots, v, extrap="flat", eval_mode="binned", **kw)
    gg = ci.convolution_interpolation(knots, v, extrap="flat", eval_mode="gather", **kw)
    assert gb.itp.subgrid == ("linear",) * 3
    o = oracle.convolution_interpolation(knots, v, extrap="flat", **kw)
    rb, rg = gb.eval(*q), gg.eval(*q)
    S = _scale(o)
    # (top-derivative tables have entries of magnitude 1e3..1e5 with alternating signs: the two strategies associate the
    # stencil sum differently, the common bound is the parity tolerance)
    assert rel_err(rb, rg, S).max() <= TOL64
    assert rel_err(rb, o.eval(*q), S).max() <= TOL64


@pytest.mark.parametrize("kw,strategy", [(dict(kernel="b5"), "cellmma"), (dict(kernel="b5", derivative=(1, 1, 1), bc="detect"), "cellmma"),
                                         (dict(kernel="b5", subgrid="linear", linear_rows=101), "cellmma"),
                                         (dict(kernel="b5", derivative=(1, 0, 2)), "cellmma"),
                                         (dict(kernel="b5", derivative=(0, 1, 0)), "cellmma"),
                                         (dict(kernel="b7"), "cellmma"), (dict(kernel="b9", derivative=(0, 0, 1)), "cellmma"),
                                         (dict(kernel="b11", bc="detect"), "cellmma_block"), (dict(kernel="b13"), "cellmma_block"),
                                         (dict(kernel="b13", derivative=(1, 1, 1)), "cellmma_block"),
                                         (dict(kernel="b13", derivative=(2, 0, 1)), "binned"),  # six 14.5-kB tables do not fit next to the brick
                                         (dict(kernel="a4"), "binned")])
def test_tensor_core_strategy_against_the_oracle(ci, oracle, kw, strategy):
    """Dense batches of a Float64 :b5-width interpolant (>= 2.5 queries per cell) take eval_cellmma_kernel: the axis-0
    contraction of up to 8 queries of an aligned cell pair as FP64 tensor-core products (DMMA), weights from the staged
    cubic-in-t tables.  Same results as the gather kernel (a few ulp: another association) and as the oracle, for plain
    evaluation, a derivative, the linear subgrid, queries outside the domain (spill bin), ragged cell-pair groups
    (1 .. 8 queries) and adversarially clustered batches; a query's bits do not depend on the batch it arrives in.
    Interpolants with more than one distinct (K, K') table pair (mixed derivative orders) and the wider kernels (:b7 .. :b13:
    12^3 .. 18^3 stencils, column tiles in chunks of 13) take the generic variant of the kernel (weights blended from the
    staged K / K' tables); the a-series stays on eval_binned_kernel."""
    import warnings
    import torch
    rng = np.random.default_rng(57)
    shape = (54, 47, 41)  # 53 cells along axis 0: the last cell pair of a row is half empty (padded extent even: TMA strides)
    knots = tuple(np.linspace(0, 1, n) for n in shape)
    v = _field(shape, rng, noise=0.02)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        gb = ci.convolution_interpolation(knots, v, extrap="flat", eval_mode="binned", **kw)
        gg = ci.convolution_interpolation(knots, v, extrap="flat", eval_mode="gather", **kw)
    o = oracle.convolution_interpolation(knots, v, extrap="flat", **{k: w for k, w in kw.items() if k != "linear_rows"}) \
        if "linear_rows" not in kw else None
    S = np.abs(v).max() if o is None else _scale(o)
    nq = 600_011  # ~6 per cell
    q = rng.uniform(-0.01, 1.01, (3, nq))
    q[:, :50_000] = (0.37 + 0.02 * rng.random((3, 50_000)))            # a dense blob: many full groups per cell pair
    q[:, 50_000:50_100] = np.array([[0.5], [0.25], [0.75]])            # 100 identical queries
    rb, rg = gb.eval(*q), gg.eval(*q)
    assert gb.itp.last_strategy() == strategy
    assert np.isfinite(rb).all()
    scale_g = S if o is not None else np.abs(rg).max()
    assert rel_err(rb, rg, scale_g).max() <= 1e-13
    if o is not None:
        assert rel_err(rb, o.eval(*q), S).max() <= TOL64
    assert np.all(rb[50_000:50_100] == rb[50_000])
    # batch independence, bitwise: a permuted batch and a sub-batch that is still dense enough for the same kernel
    perm = rng.permutation(nq)
    rp = gb.eval(*q[:, perm])
    assert np.array_equal(rp, rb[perm])
    # sub-batches: sparser batches take the rows of a product from 2 x 2 x 2 cell blocks instead of cell pairs (another
    # association: a few ulp); the bits of EITHER kernel do not depend on the batch a query arrives in
    seen = {strategy: rb}
    for frac in (2, 3, 5):
        n = nq // frac
        r = gb.eval(*np.ascontiguousarray(q[:, :n]))
        st = gb.itp.last_strategy()
        assert rel_err(r, rb[:n], scale_g).max() <= 1e-13, (frac, st)
        if st in seen:
            assert np.array_equal(r, seen[st][:n]), (frac, st)
        else:
            seen[st] = r
            if o is not None:
                assert rel_err(r, o.eval(*q[:, :n]), S).max() <= TOL64
    if strategy.startswith("cellmma"):
        assert "cellmma_block" in seen
    # device-resident tensors through the same path
    tq = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in q]
    assert np.array_equal(gb.eval(*tq).cpu().numpy(), rb)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_interpolants_with_several_table_pairs_at_one_query_per_cell(ci, oracle, dtype):
    """A :b5 interpolant with mixed derivative orders at 1.2 queries per cell: Float64 takes the generic tensor-core kernel
    on 2 x 2 x 2 cell blocks (weights blended from the staged K / K' tables), Float32 - no generic tensor-core kernel there -
    eval_cellrun_kernel (cell sort + register-resident stencils, position records); both against the oracle."""
    import torch
    rng = np.random.default_rng(61)
    shape = (56, 64, 48)
    knots = tuple((np.arange(n) / 64).astype(dtype) for n in shape)
    v = _field(shape, rng, noise=0.01).astype(dtype)
    kw = dict(kernel="b5", derivative=(1, 0, 0), bc="poly")
    itp = ci.convolution_interpolation(knots, v, **kw)
    o = oracle.convolution_interpolation(tuple(k.astype(np.float64) for k in knots), v.astype(np.float64), **kw)
    nq = 200_003  # 1.2 per cell
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    xs = [(torch.rand(nq, dtype=tdt, device="cuda") * float(k[-1])).clamp(0.0, float(k[-1])) for k in knots]
    out = itp.eval(*xs)
    assert itp.itp.last_strategy() == ("cellmma_block" if dtype == np.float64 else "cellrun")
    ref = o.eval(*[x.cpu().numpy().astype(np.float64) for x in xs])
    assert rel_err(out.cpu().numpy(), ref, _scale(o)).max() <= (TOL64 if dtype == np.float64 else TOL32)


def test_device_tensors_check_finite_option(ci):
    """CUDA tensors with extrap=Flat: a NaN coordinate yields NaN without a host synchronisation, and raises InexactError
    like the host-buffer path when the caller asks for the status word (check_finite=True)."""
    import torch
    rng = np.random.default_rng(3)
    k = tuple(np.linspace(0, 1, 24) for _ in range(3))
    itp = ci.convolution_interpolation(k, _field((24, 24, 24), rng), extrap="flat")
    q = [torch.rand(5000, dtype=torch.float64, device="cuda") for _ in range(3)]
    q[1][1234] = float("nan")
    out = itp.eval(*q)
    assert bool(torch.isnan(out[1234])) and int(torch.isnan(out).sum()) == 1
    with pytest.raises(ci.InexactError, match="query 1234 "):
        itp.eval(*q, check_finite=True)


def test_last_strategy_reports_the_kernel_of_the_last_call(ci):
    """cib200_itp_last_strategy: AUTO on a 128^3 :b5 grid takes the gather kernel for a small batch, the binned kernel for
    a sparse large one, the tensor-core kernel on 2 x 2 x 2 cell blocks from 0.4 query per cell and on cell pairs from 3."""
    import torch
    rng = np.random.default_rng(12)
    n = 128
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    itp = ci.convolution_interpolation(k, _field((n, n, n), rng), bc="poly")
    ncell = (n - 1) ** 3
    for nq, want in ((10_000, "gather"), (int(0.2 * ncell), "binned"), (int(1.0 * ncell), "cellmma_block"), (int(2.5 * ncell), "cellmma_block"),
                     (int(4.0 * ncell), "cellmma")):
        q = [torch.rand(nq, dtype=torch.float64, device="cuda") for _ in range(3)]
        itp.eval(*q)
        assert itp.itp.last_strategy() == want, (nq, want, itp.itp.last_strategy())


def test_binned_device_resident_and_auto_mode(ci):
    import torch
    rng = np.random.default_rng(32)
    n = 128
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    v = _field((n, n, n), rng)
    q = [torch.rand(5_000_000, dtype=torch.float64, device="cuda") for _ in range(3)]
    ga = ci.convolution_interpolation(k, v, bc="poly")                       # auto -> binned for >= 2^22 queries
    gg = ci.convolution_interpolation(k, v, bc="poly", eval_mode="gather")
    ra, rg = ga.eval(*q), gg.eval(*q)
    assert float((ra - rg).abs().max()) <= 1e-14 * float(np.abs(v).max())
    assert ga.itp.eval_mode == "auto"
    perm = torch.randperm(5_000_000, device="cuda")
    assert torch.equal(ga.eval(*[x[perm] for x in q]), ra[perm])            # permutation invariance, bitwise


def test_binned_clustered_query_distributions(ci):
    """Adversarial batches for the binning passes (everything in one cell / one tile / one plane of cells / a narrow
    blob): the binned strategies must return what the gather strategy returns."""
    import torch
    rng = np.random.default_rng(34)
    n = 96
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    v = _field((n, n, n), rng)
    gb = ci.convolution_interpolation(k, v, bc="poly", eval_mode="binned")
    gg = ci.convolution_interpolation(k, v, bc="poly", eval_mode="gather")
    nq, h = 3_000_000, 1.0 / (n - 1)
    u = lambda: torch.rand(nq, dtype=torch.float64, device="cuda")
    blob = lambda: (0.5 + 0.01 * torch.randn(nq, dtype=torch.float64, device="cuda")).clamp_(0, 1)
    cases = {"one cell": [0.5 + h * u(), 0.25 + h * u(), 0.75 + h * u()],
             "one tile": [0.5 + 16 * h * u(), 0.25 + 16 * h * u(), 0.75 + 16 * h * u()],
             "one plane": [u(), u(), 0.5 + h * u()],
             "blob": [blob(), blob(), blob()]}
    for name, q in cases.items():
        rb, rg = gb.eval(*q), gg.eval(*q)
        assert float((rb - rg).abs().max()) <= 1e-13 * float(np.abs(v).max()), name


def test_host_buffer_pipeline_equals_device_resident(ci):
    """cib200_itp_eval with HOST buffers runs a chunked H2D -> kernels -> D2H pipeline (2^24-query chunks, ring of 3
    device buffers, one compute stream).  Several chunks in flight must give the same bits as one device-resident
    call, in both strategies, including a ragged last chunk."""
    import torch
    rng = np.random.default_rng(33)
    n = 96
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    v = _field((n, n, n), rng)
    nq = 3 * (1 << 24) + 12_345
    q = rng.random((3, nq))
    tq = [torch.from_numpy(a).cuda() for a in q]
    for mode in ("binned", "gather"):
        g = ci.convolution_interpolation(k, v, bc="poly", eval_mode=mode)
        dev = g.eval(*tq).cpu().numpy()
        host = g.eval(*q)
        # (binned: the 2^24-query chunks take the tensor-core kernel like the device-resident call, the ragged last chunk
        # of 12 345 queries the sparse kernel - another association of the stencil sum, a few ulp)
        full = 3 * (1 << 24)
        assert np.array_equal(host[:full], dev[:full]), mode
        assert np.abs(host[full:] - dev[full:]).max() <= (0.0 if mode == "gather" else 1e-14 * np.abs(v).max()), mode
        g.itp.close()


@pytest.mark.parametrize("kernel,bc", [("b5", "detect"), ("b13", "poly"), ("a4", "detect"), ("b7", "linear"),
                                       ("b9", ("quadratic", "detect"))])
def test_device_side_coefficient_build_equals_host_build(ci, oracle, kernel, bc):
    """cib200_build_coefs_device against the host build and the oracle's create_coefs: same ghost
    branches (:detect decisions included), values equal up to the association of the line mean"""
    rng = np.random.default_rng(77)
    for shape in ((41,), (23, 37), (19, 33, 26), (9, 12, 10, 11)):
        k = tuple(np.linspace(0, 1 + 0.2 * i, s) for i, s in enumerate(shape))
        v = _field(shape, rng, noise=0.3 if bc == "detect" else 0.05)  # noisy lines exercise both :detect outcomes
        kk = k[0] if len(shape) == 1 else k
        gh = ci.convolution_interpolation(kk, v, kernel=kernel, bc=bc, extrap="flat")
        gd = ci.convolution_interpolation(kk, v, kernel=kernel, bc=bc, extrap="flat", coefs_on="device")
        o = oracle.convolution_interpolation(kk, v, kernel=kernel, bc=bc, extrap="flat")
        ch, cd = gh.itp.coefs, gd.itp.coefs
        assert ch.shape == cd.shape == o.coefs.shape
        scale = np.abs(o.coefs).max()
        assert np.abs(cd - o.coefs).max() <= 1e-13 * scale
        assert np.abs(cd - ch).max() <= 1e-13 * scale
        q = np.stack([rng.uniform(-0.05, kk_[-1] + 0.05, 2000) for kk_ in k])
        assert rel_err(gd.eval(*q), o.eval(*q), _scale(o)).max() <= TOL64


def test_device_side_coefficient_build_float32_and_c_abi_host_buffers(ci, oracle):
    import ctypes as C
    rng = np.random.default_rng(78)
    shape = (30, 28, 31)
    v = _field(shape, rng).astype(np.float32)
    k = tuple(np.linspace(0, 1, s) for s in shape)
    gh = ci.convolution_interpolation(k, v, kernel="b5", bc="poly")
    gd = ci.convolution_interpolation(k, v, kernel="b5", bc="poly", coefs_on="device")
    assert gd.itp.coefs.dtype == np.float32
    assert np.abs(gd.itp.coefs.astype(np.float64) - gh.itp.coefs).max() <= 2e-6 * np.abs(gh.itp.coefs).max()
    # the C entry point with HOST buffers (staged through the device)
    L = ci.lib()
    v64 = np.asfortranarray(_field(shape, rng))
    out = np.zeros(tuple(s + 8 for s in shape), order="F")
    n = (C.c_int64 * 3)(*shape)
    hh = (C.c_double * 3)(*[kk[1] - kk[0] for kk in k])
    kid = (C.c_int32 * 3)(0, 0, 0)
    bc = (C.c_int32 * 3)(0, 0, 0)
    ci.check(L.cib200_build_coefs_device(3, 0, n, v64.ctypes.data, hh, kid, bc, bc, out.ctypes.data, 0, None))
    o = oracle.convolution_interpolation(k, v64, kernel="b5", bc="detect")
    assert np.abs(out - o.coefs).max() <= 1e-13 * np.abs(o.coefs).max()


def test_concurrent_calls_on_one_handle_are_serialised(ci):
    """Threads.@threads over chunks of a query vector is the reference's threading pattern: concurrent calls on one
    handle (shared binning scratch / staging ring) must give the same bits as sequential calls"""
    import threading
    import torch
    rng = np.random.default_rng(91)
    k = tuple(np.linspace(0, 1, 48) for _ in range(3))
    g = ci.convolution_interpolation(k, _field((48, 48, 48), rng), extrap="flat", eval_mode="binned")
    chunks = [rng.uniform(-0.02, 1.02, (3, 300_000)) for _ in range(6)]
    seq = [g.eval(*c) for c in chunks]
    res = [None] * len(chunks)

    def work(i):
        res[i] = g.eval(*chunks[i])

    ts = [threading.Thread(target=work, args=(i,)) for i in range(len(chunks))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for a, b in zip(seq, res):
        assert np.array_equal(a, b)
    # device-resident calls on different streams
    dq = [[torch.from_numpy(a).cuda() for a in c] for c in chunks]
    streams = [torch.cuda.Stream() for _ in chunks]
    outs = []
    torch.cuda.synchronize()
    for s_, q in zip(streams, dq):
        with torch.cuda.stream(s_):
            outs.append(g.eval(*q))
    torch.cuda.synchronize()
    for a, o in zip(seq, outs):
        assert np.array_equal(a, o.cpu().numpy())


def test_single_process_multi_gpu_evaluation(ci, oracle):
    """cib200_itp_eval_multi: one handle per device, contiguous query slices, lowest offending index of the whole batch.
    With one GPU the two handles share device 0 (the split / merge logic is the same)."""
    import torch
    ndev = torch.cuda.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    rng = np.random.default_rng(92)
    k = tuple(np.linspace(0, 1, 40) for _ in range(3))
    v = _field((40, 40, 40), rng)
    one = ci.convolution_interpolation(k, v, extrap="line")
    multi = ci.convolution_interpolation(k, v, extrap="line", devices=devices)
    q = rng.uniform(-0.03, 1.03, (3, 1_000_003))
    a, b = one.eval(*q), multi.eval(*q)
    assert np.array_equal(a, b)
    o = oracle.convolution_interpolation(k, v, extrap="line")
    sub = slice(0, 200_000)
    qs = tuple(x[sub] for x in q)
    assert rel_err(b[sub], o.eval(*qs), _scale(o) * _line_amplification(o, qs)).max() <= TOL64
    thr = ci.convolution_interpolation(k, v, devices=devices)
    qq = np.full((3, 400_000), 0.5)
    qq[1, 350_000] = 1.5   # second slice
    qq[0, 399_999] = np.nan
    with pytest.raises(ci.DomainError, match="absolute query index 350000"):
        thr.eval(*qq)
    qq[2, 10] = -0.5       # first slice: the lowest index wins
    with pytest.raises(ci.DomainError, match="absolute query index 10"):
        thr.eval(*qq)
    # 2-D and tiny batches
    k2 = (np.linspace(0, 1, 30), np.linspace(0, 2, 35))
    v2 = rng.random((30, 35))
    m2 = ci.convolution_interpolation(k2, v2, derivative=(1, 0), extrap="flat", devices=devices)
    s2 = ci.convolution_interpolation(k2, v2, derivative=(1, 0), extrap="flat")
    for n in (1, 3, 1000):
        q2 = (rng.uniform(0, 1, n), rng.uniform(0, 2, n))
        assert np.array_equal(m2.eval(*q2), s2.eval(*q2))
