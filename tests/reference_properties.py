"""Ports of the analytic assertions of the reference's own test-suite (/root/reference/test/*.jl) for the
hot path.  Each check takes an `impl` namespace with

    impl.interp(knots, values, **kw) -> object with .eval(*coords) (element-wise) and __call__ (Julia call)
    impl.resample(knots_in, knots_out, values, **kw)
    impl.smooth(knots, values, B)

so the same assertions run against the CPU oracle (no GPU) and against the CUDA library (-m gpu).
Unseeded rand() of the reference is replaced by seeded generators.
"""
import numpy as np

TWO_PI = 2 * np.pi


def check_grid_point_reproduction(impl):
    """test_uniform_interpolation.jl:31-49,72-91,120-145 - random data reproduced at the grid points, 1e-6"""
    rng = np.random.default_rng(1)
    N = 12
    for kernel in ("b5", "b13"):
        r = np.linspace(0.0, 1.0, N)
        v1 = rng.random(N)
        itp = impl.interp(r, v1, kernel=kernel, bc="linear")
        assert np.abs(itp.eval(r) - v1).max() < 1e-6
        v2 = rng.random((N, N))
        itp = impl.interp((r, r), v2, kernel=kernel, bc="linear")
        got = np.array([[itp(float(a), float(b)) for b in r] for a in r])
        assert np.abs(got - v2).max() < 1e-6
        assert np.abs(itp(r, r) - v2).max() < 1e-6  # vector call = Cartesian product
        v3 = rng.random((N, N, N))
        itp = impl.interp((r, r, r), v3, kernel=kernel, bc="linear")
        assert np.abs(itp(r, r, r) - v3).max() < 1e-6


def check_multilinear_midpoints(impl):
    """test_uniform_interpolation.jl:194-239 - multilinear data reproduced at cell midpoints, 1e-6"""
    N = 8
    r = np.linspace(0.0, 1.0, N)
    mid = 0.5 * (r[1:] + r[:-1])
    itp = impl.interp(r, r.copy(), kernel="b5", bc="linear")
    assert np.abs(itp.eval(mid) - mid).max() < 1e-6
    v2 = r[:, None] + 2 * r[None, :]
    itp = impl.interp((r, r), v2, kernel="b5", bc="linear")
    assert np.abs(itp(mid, mid) - (mid[:, None] + 2 * mid[None, :])).max() < 1e-6
    v3 = r[:, None, None] + 2 * r[None, :, None] - r[None, None, :]
    itp = impl.interp((r, r, r), v3, kernel="b5", bc="linear")
    ref = mid[:, None, None] + 2 * mid[None, :, None] - mid[None, None, :]
    assert np.abs(itp(mid, mid, mid) - ref).max() < 1e-6


def check_convergence_order(impl):
    """test_uniform_convergence.jl:19-127 - b5: error ratio per grid doubling > 2^6 (d0), 2^5 (d1), 2^4 (d2)"""
    fs = [np.sin, np.cos, lambda z: -np.sin(z)]
    for d, bound in ((0, 2 ** 6), (1, 2 ** 5), (2, 2 ** 4)):
        errs = []
        for n in (12, 24, 48):
            x = np.linspace(0, TWO_PI, n)
            itp = impl.interp(x, np.sin(x), kernel="b5", bc="poly", derivative=d)
            xq = np.linspace(0, TWO_PI, 997)
            errs.append(np.abs(itp.eval(xq) - fs[d](xq)).max())
        assert errs[0] / errs[1] > bound and errs[1] / errs[2] > bound, (d, errs)


def check_uniform_derivatives(impl):
    """test_uniform_derivatives.jl - derivative=1 (mixed partial in N-D) of sin products, N=20, bc=:poly, 1e-2"""
    N = 20
    x = np.linspace(0, TWO_PI, N)
    q = np.linspace(0.3, 6.0, 23)
    itp = impl.interp(x, np.sin(x), kernel="b5", bc="poly", derivative=1)
    assert np.abs(itp.eval(q) - np.cos(q)).max() < 1e-2
    v2 = np.sin(x)[:, None] * np.sin(x)[None, :]
    itp = impl.interp((x, x), v2, kernel="b5", bc="poly", derivative=1)
    assert np.abs(itp(q, q) - np.cos(q)[:, None] * np.cos(q)[None, :]).max() < 1e-2
    v3 = v2[:, :, None] * np.sin(x)[None, None, :]
    itp = impl.interp((x, x, x), v3, kernel="b5", bc="poly", derivative=1)
    ref = np.cos(q)[:, None, None] * np.cos(q)[None, :, None] * np.cos(q)[None, None, :]
    assert np.abs(itp(q, q, q) - ref).max() < 1e-2


def check_perdim_derivatives(impl):
    """test_perdim_derivatives.jl:43-141 - per-axis derivative tuples, N=50, 1e-4"""
    N = 50
    x = np.linspace(0, TWO_PI, N)
    q = np.linspace(0.4, 5.9, 17)
    s, c = np.sin, np.cos
    f2 = s(x)[:, None] * c(x)[None, :]
    cases2 = {(1, 0): lambda a, b: c(a) * c(b), (0, 1): lambda a, b: -s(a) * s(b),
              (1, 1): lambda a, b: -c(a) * s(b), (2, 0): lambda a, b: -s(a) * c(b),
              (0, 2): lambda a, b: -s(a) * c(b)}
    for dv, f in cases2.items():
        itp = impl.interp((x, x), f2, kernel="b5", bc="poly", derivative=dv)
        ref = f(q[:, None], q[None, :])
        assert np.abs(itp(q, q) - ref).max() < 1e-4, dv
    f3 = s(x)[:, None, None] * c(x)[None, :, None] * s(x)[None, None, :]
    A, B, Cc = q[:, None, None], q[None, :, None], q[None, None, :]
    cases3 = {(1, 0, 0): c(A) * c(B) * s(Cc), (0, 1, 0): -s(A) * s(B) * s(Cc), (0, 0, 1): s(A) * c(B) * c(Cc),
              (1, 1, 0): -c(A) * s(B) * s(Cc), (0, 1, 1): -s(A) * s(B) * c(Cc)}
    for dv, ref in cases3.items():
        itp = impl.interp((x, x, x), f3, kernel="b5", bc="poly", derivative=dv)
        assert np.abs(itp(q, q, q) - ref).max() < 1e-4, dv


def check_extrapolation(impl):
    """test_extrapolation.jl:26-38,55-68,88-103 - :line continues f = sum x_d linearly, :flat clamps (N=5, bc=:linear)"""
    N = 5
    r = np.linspace(0.0, 1.0, N)
    pts = np.array([-0.2, -0.1, 1.1, 1.2])
    etp = impl.interp(r, r.copy(), kernel="b5", bc="linear", extrap="line")
    assert np.abs(etp.eval(pts) - pts).max() < 1e-6
    etp = impl.interp(r, r.copy(), kernel="b5", bc="linear", extrap="flat")
    assert np.abs(etp.eval(pts) - np.array([0.0, 0.0, 1.0, 1.0])).max() < 1e-6
    v2 = r[:, None] + r[None, :]
    etp = impl.interp((r, r), v2, kernel="b5", bc="linear", extrap="line")
    assert np.abs(np.array([etp(float(i), float(i)) for i in pts]) - 2 * pts).max() < 1e-6
    etp = impl.interp((r, r), v2, kernel="b5", bc="linear", extrap="flat")
    assert np.abs(np.array([etp(float(i), float(i)) for i in pts]) - np.array([0, 0, 2.0, 2.0])).max() < 1e-6
    v3 = r[:, None, None] + r[None, :, None] + r[None, None, :]
    etp = impl.interp((r, r, r), v3, kernel="b5", bc="linear", extrap="line")
    assert np.abs(np.array([etp(float(i), float(i), float(i)) for i in pts]) - 3 * pts).max() < 1e-3
    etp = impl.interp((r, r, r), v3, kernel="b5", bc="linear", extrap="flat")
    assert np.abs(np.array([etp(float(i), float(i), float(i)) for i in pts]) - np.array([0, 0, 3.0, 3.0])).max() < 1e-6


def check_extrapolation_4d(impl):
    """test_extrapolation.jl:124-140 (4D fast kernels, N=5, bc=:linear): the HigherDimension evaluators (eager_nd.jl)
    behind :line / :flat; plus an analytic derivative through the per-dim N-D evaluator (perdim_nd.jl)"""
    N = 5
    r = np.linspace(0.0, 1.0, N)
    pts = np.array([-0.2, -0.1, 1.1, 1.2])
    v4 = r[:, None, None, None] + r[None, :, None, None] + r[None, None, :, None] + r[None, None, None, :]
    for kernel in ("a3", "b5"):
        etp = impl.interp((r, r, r, r), v4, kernel=kernel, bc="linear", extrap="line")
        assert np.abs(etp.eval(pts, pts, pts, pts) - 4 * pts).max() < 1e-3
        etp = impl.interp((r, r, r, r), v4, kernel=kernel, bc="linear", extrap="flat")
        assert np.abs(etp.eval(pts, pts, pts, pts) - np.array([0, 0, 4.0, 4.0])).max() < 1e-6
    n = 20
    x = np.linspace(0, TWO_PI, n)
    g = np.meshgrid(x, x, x, x, indexing="ij")
    v = np.sin(g[0]) * np.cos(g[1]) * np.sin(g[2]) * np.cos(g[3])
    q = np.random.default_rng(7).uniform(0.5, TWO_PI - 0.5, (4, 300))
    itp = impl.interp((x, x, x, x), v, kernel="b5", derivative=(1, 0, 0, 0))
    ref = np.cos(q[0]) * np.cos(q[1]) * np.sin(q[2]) * np.cos(q[3])
    assert np.abs(itp.eval(*q) - ref).max() < 1e-3
    itp = impl.interp((x, x, x, x), v, kernel="b5")
    ref = np.sin(q[0]) * np.cos(q[1]) * np.sin(q[2]) * np.cos(q[3])
    assert np.abs(itp.eval(*q) - ref).max() < 1e-4


def check_resample(impl):
    """test_resample.jl:4-80 - resample of sin products vs analytic (1e-5 / 1e-4 / 1e-4), sizes, derivative"""
    xi, xo = np.linspace(0, TWO_PI, 50), np.linspace(0, TWO_PI, 100)
    out = impl.resample(xi, xo, np.sin(xi))
    assert out.shape == (100,)
    assert np.abs(out - np.sin(xo)).max() < 1e-5
    xi2, xo2 = np.linspace(0, TWO_PI, 50), np.linspace(0, TWO_PI, 60)
    v2 = np.sin(xi2)[:, None] * np.cos(xi2)[None, :]
    out = impl.resample((xi2, xi2), (xo2, xo2), v2)
    assert out.shape == (60, 60)
    assert np.abs(out - np.sin(xo2)[:, None] * np.cos(xo2)[None, :]).max() < 1e-4
    xi3, xo3 = np.linspace(0, TWO_PI, 40), np.linspace(0, TWO_PI, 60)
    v3 = np.sin(xi3)[:, None, None] * np.cos(xi3)[None, :, None] * np.sin(xi3)[None, None, :]
    out = impl.resample((xi3,) * 3, (xo3,) * 3, v3, kernel="b7")
    assert out.shape == (60, 60, 60)
    ref3 = np.sin(xo3)[:, None, None] * np.cos(xo3)[None, :, None] * np.sin(xo3)[None, None, :]
    assert np.abs(out - ref3).max() < 1e-4
    # accuracy ordering b5 >= b13 error (test_resample.jl: a1 >= a3 >= a4 >= b5)
    e5 = np.abs(impl.resample(xi, xo, np.sin(xi), kernel="b5") - np.sin(xo)).max()
    e13 = np.abs(impl.resample(xi, xo, np.sin(xi), kernel="b13") - np.sin(xo)).max()
    assert e13 <= e5 * 1.0000001
    # derivative=1 with b7, 1e-3
    d = impl.resample((xi,), (xo,), np.sin(xi), kernel="b7", derivative=1)
    assert np.abs(d - np.cos(xo)).max() < 1e-3
    # non-matching output sizes per axis
    out = impl.resample((xi2, xi2), (xo2, xo), v2, kernel="b5")
    assert out.shape == (60, 100)


def check_smooth(impl):
    """test_gaussian.jl:22-84 - smooth recovers sin/cos within 0.15, size/eltype preserved, roughness ordering"""
    rng = np.random.default_rng(7)
    xs = np.linspace(0, TWO_PI, 200)
    noisy = np.sin(xs) + 0.1 * rng.standard_normal(200)
    sm = impl.smooth(xs, noisy, 0.05)
    assert sm.shape == (200,) and sm.dtype == np.float64
    assert np.abs(sm - np.sin(xs)).max() < 0.15
    assert np.abs(impl.smooth(xs, np.cos(xs), 0.05) - np.cos(xs)).max() < 0.15
    x2 = np.linspace(0, TWO_PI, 100)
    sig2 = np.sin(x2)[:, None] * np.cos(x2)[None, :]
    sm2 = impl.smooth((x2, x2), sig2 + 0.1 * rng.standard_normal((100, 100)), 0.05)
    assert sm2.shape == (100, 100)
    assert np.abs(sm2 - sig2).max() < 0.25
    rough = lambda v: np.abs(np.diff(v, 2)).sum()
    assert rough(impl.smooth(xs, noisy, 0.02)) < rough(impl.smooth(xs, noisy, 0.1)) < rough(noisy)


ALL_CHECKS = [check_grid_point_reproduction, check_multilinear_midpoints, check_convergence_order,
              check_uniform_derivatives, check_perdim_derivatives, check_extrapolation, check_extrapolation_4d,
              check_resample, check_smooth]
