"""Parity of the CUDA axis sweeps (cib200_resample / cib200_smooth through the C-ABI) with the CPU oracle.

Tolerance: <= 1e-12 * max(|ref|, S), S = max|values| * prod_d (1/h_d)^DO_d (Float64); <= 1e-5 (Float32).
"""
import numpy as np
import pytest

from .conftest import rel_err

pytestmark = pytest.mark.gpu
TOL64 = 1e-12
TWO_PI = 2 * np.pi


def _sin_field(shape, rng=None, noise=0.0):
    grids = np.meshgrid(*[np.linspace(0, TWO_PI, n) for n in shape], indexing="ij")
    f = np.ones(shape)
    for i, g in enumerate(grids):
        f = f * (np.sin(g) if i % 2 == 0 else np.cos(g))
    if rng is not None and noise:
        f = f + noise * rng.standard_normal(shape)
    return f


def _extrapolation_conditioning(kin, kout):
    """Output knots outside the input range are not rejected: the edge cell's table row is Hermite-EXTRAPOLATED
    with |t| >> 1 (convolution_resample.jl:322-333 clamps i and idx but not t).  The basis polynomials then grow
    like |t|^3 with massive cancellation, so ANY two evaluation orders differ by ~eps |t|^3; the parity scale is
    multiplied by (1 + |t|)^3 for those knots.  In-range knots keep the plain 1e-12."""
    amps = []
    for ki, ko in zip(kin, kout):
        h = ki[1] - ki[0]
        outside = np.maximum(np.maximum(ki[0] - ko, ko - ki[-1]), 0.0) / h
        amps.append((1.0 + 100.0 * outside) ** 3)
    return np.prod(np.meshgrid(*amps, indexing="ij"), axis=0) if len(amps) > 1 else amps[0]


def _check_resample(ci, oracle, kin, kout, v, tol=TOL64, **kw):
    ref = oracle.convolution_resample(kin, kout, v, **kw)
    got = ci.convolution_resample(kin, kout, v, **kw)
    assert got.shape == ref.shape
    nd = v.ndim
    der = kw.get("derivative", 0)
    der = (der,) * nd if isinstance(der, int) else tuple(der)
    kin_t = kin if isinstance(kin, tuple) else (kin,)
    S = np.abs(v).max()
    for k, d in zip(kin_t, der):
        S *= (1.0 / (k[1] - k[0])) ** d
    kout_t = kout if isinstance(kout, tuple) else (kout,)
    err = rel_err(got, ref, S * _extrapolation_conditioning(kin_t, kout_t))
    assert err.max() <= tol, err.max()
    return got, ref


def test_c3_resample_b13_128_to_256_with_derivative_pass(ci, oracle):
    """BASELINE config 3 at full size: 3-D :b13 128^3 -> 256^3, plus the derivative=(1,0,0) pass."""
    rng = np.random.default_rng(0xC1B200 + 3)
    kin = tuple(np.linspace(0, TWO_PI, 128) for _ in range(3))
    kout = tuple(np.linspace(0, TWO_PI, 256) for _ in range(3))
    v = _sin_field((128, 128, 128))
    got, _ = _check_resample(ci, oracle, kin, kout, v, kernel="b13")
    xo = kout[0]
    exact = np.sin(xo)[:, None, None] * np.cos(xo)[None, :, None] * np.sin(xo)[None, None, :]
    assert np.abs(got - exact).max() < 1e-9
    vn = v + 0.02 * rng.standard_normal(v.shape)  # noise flips :detect to the linear fallback on many lines
    _check_resample(ci, oracle, kin, kout, vn, kernel="b13", derivative=(1, 0, 0))


@pytest.mark.parametrize("kernel", ["b5", "b7", "b9", "b11", "b13", "a3", "a4", "a5", "a7"])
@pytest.mark.parametrize("bc", ["detect", "poly", "linear", "quadratic"])
def test_resample_1d_2d_kernels_and_bcs(ci, oracle, kernel, bc):
    rng = np.random.default_rng(21)
    xi, xo = np.linspace(0, TWO_PI, 50), np.linspace(0, TWO_PI, 100)
    _check_resample(ci, oracle, xi, xo, np.sin(xi) + 0.01 * rng.standard_normal(50), kernel=kernel, bc=bc)
    k2i = (np.linspace(0, 1, 37), np.linspace(-1, 1, 29))
    k2o = (np.linspace(0, 1, 64), np.linspace(-1, 1, 45))
    _check_resample(ci, oracle, k2i, k2o, rng.random((37, 29)), kernel=kernel, bc=bc,
                    derivative=(0, 1) if kernel.startswith("b") else 0)
    if kernel.startswith("a") and bc == "detect":  # 3-D, through the pipelined passes
        k3i = tuple(np.linspace(0, 1, n) for n in (40, 33, 36))
        k3o = tuple(np.linspace(0, 1, n) for n in (64, 70, 41))
        _check_resample(ci, oracle, k3i, k3o, rng.random((40, 33, 36)), kernel=kernel)


def test_resample_quintic_subgrid(ci, oracle):
    rng = np.random.default_rng(26)
    xi, xo = np.linspace(0, TWO_PI, 60), np.linspace(0, TWO_PI, 131)
    v = np.sin(xi) + 0.01 * rng.standard_normal(60)
    _check_resample(ci, oracle, xi, xo, v, kernel="b5", subgrid="quintic")
    _check_resample(ci, oracle, xi, xo, v, kernel="b13", subgrid="quintic", derivative=2)
    k2i = (np.linspace(0, 1, 37), np.linspace(-1, 1, 29))
    k2o = (np.linspace(0, 1, 64), np.linspace(-1, 1, 45))
    _check_resample(ci, oracle, k2i, k2o, rng.random((37, 29)), kernel=("b7", "b5"), subgrid=("quintic", "cubic"),
                    derivative=(1, 0))
    with pytest.raises(ci.UnsupportedModeError):
        ci.convolution_resample(xi, xo, v, kernel="b5", subgrid="quintic", derivative=2)  # K^(4) of b5 not shipped
    # :linear subgrid: explicit, through a top derivative, and the 3-D kw method (resolves to all-linear)
    _check_resample(ci, oracle, xi, xo, v, kernel="b5", subgrid="linear")
    _check_resample(ci, oracle, xi, xo, v, kernel="b5", derivative=3)
    k3i = tuple(np.linspace(0, 1, n) for n in (30, 28, 33))
    k3o = tuple(np.linspace(0, 1, n) for n in (40, 50, 61))
    _check_resample(ci, oracle, k3i, k3o, rng.random((30, 28, 33)), kernel="b7", derivative=(4, 0, 1))


def test_resample_shapes_downsampling_extrapolating_knots_and_4d(ci, oracle):
    rng = np.random.default_rng(22)
    # output knots outside the input range are silently clamped to the edge cell (convolution_resample.jl:322)
    xi = np.linspace(0, 1, 40)
    xo = np.linspace(-0.2, 1.3, 77)
    _check_resample(ci, oracle, xi, xo, rng.random(40), kernel="b13")
    # downsampling, non-uniform output knots, mixed kernels / derivatives / bcs per axis
    k3i = (np.linspace(0, 1, 30), np.linspace(0, 2, 41), np.linspace(0, 3, 25))
    k3o = (np.sort(rng.uniform(0, 1, 13)), np.linspace(0, 2, 90), np.linspace(0.5, 2.5, 7))
    _check_resample(ci, oracle, k3i, k3o, rng.random((30, 41, 25)), kernel=("b5", "b13", "b9"),
                    derivative=(1, 0, 2), bc=(("poly", "linear"), ("detect", "detect"), ("quadratic", "poly")))
    k4i = tuple(np.linspace(0, 1, n) for n in (12, 11, 13, 10))
    k4o = tuple(np.linspace(0, 1, n) for n in (20, 9, 15, 17))
    # N = 4: only the kw-less call keeps the cubic subgrid (convolution_resample.jl:65-172); with keyword
    # arguments _build_subgrids turns N >= 4 into :linear (K rows only, linear blend)
    v4 = rng.random((12, 11, 13, 10))
    got_c, _ = _check_resample(ci, oracle, k4i, k4o, v4)
    got_l, _ = _check_resample(ci, oracle, k4i, k4o, v4, kernel="b13")
    assert 0 < np.abs(got_c - got_l).max() < 1e-3  # a 101-row table blended linearly is ~1e-4 accurate
    # a long 1-D line (single line -> tile width 1)
    xl = np.linspace(0, 50, 9001)
    _check_resample(ci, oracle, xl, np.linspace(0, 50, 12345), np.sin(xl) + 0.001 * rng.standard_normal(9001), kernel="b13")


def _check_smooth(ci, oracle, knots, v, B, tol=TOL64):
    ref = oracle.convolution_smooth(knots, v, B)
    got = ci.convolution_smooth(knots, v, B)
    assert got.shape == ref.shape == v.shape
    err = rel_err(got, ref, np.abs(v).max())
    assert err.max() <= tol, err.max()
    return got


def test_c4_smooth_4d_40_and_64(ci, oracle):
    """BASELINE config 4 at full size: 4-D Gaussian smoothing, B = 0.05 (E = 24, 49 taps), 40^4 and 64^4."""
    rng = np.random.default_rng(0xC1B200 + 4)
    for n in (40, 64):
        knots = tuple(np.linspace(0, TWO_PI, n) for _ in range(4))
        v = _sin_field((n,) * 4, rng, 0.1)
        _check_smooth(ci, oracle, knots, v, 0.05)
    v5 = _sin_field((40,) * 4) + 5.0  # the +5 variant exposes the centred-ghost quirk
    _check_smooth(ci, oracle, tuple(np.linspace(0, TWO_PI, 40) for _ in range(4)), v5, 0.05)


@pytest.mark.parametrize("B", [0.02, 0.05, 0.1, 0.2, 0.3, 0.5, 1.0])  # 77 (run-time loop), 49, 35, 25, 21 (loop), 17, 13 taps
def test_smooth_1d_2d_3d(ci, oracle, B):
    rng = np.random.default_rng(23)
    x = np.linspace(0, TWO_PI, 200)
    _check_smooth(ci, oracle, x, np.cos(x) + 0.1 * rng.standard_normal(200), B)
    _check_smooth(ci, oracle, (x[:100], x[:120]), rng.random((100, 120)), B)
    _check_smooth(ci, oracle, (x[:50], x[:45], x[:60]), rng.random((50, 45, 60)) + 3.0, B)


def test_sweeps_device_resident_and_float32(ci, oracle):
    import torch
    rng = np.random.default_rng(24)
    kin = tuple(np.linspace(0, 1, n) for n in (33, 40, 28))
    kout = tuple(np.linspace(0, 1, n) for n in (50, 41, 60))
    v = rng.random((33, 40, 28))
    host = ci.convolution_resample(kin, kout, v, kernel="b13")
    dev = ci.convolution_resample(kin, kout, torch.from_numpy(v).cuda(), kernel="b13")
    assert dev.is_cuda and tuple(dev.shape) == (50, 41, 60)
    assert np.array_equal(dev.cpu().numpy(), host)
    ks = tuple(np.linspace(0, 1, n) for n in (60, 70))
    v2 = rng.random((60, 70))
    assert np.array_equal(ci.convolution_smooth(ks, torch.from_numpy(v2).cuda(), 0.05).cpu().numpy(),
                          ci.convolution_smooth(ks, v2, 0.05))
    # Float32 opt-in: <= 1e-5 against the Float64 result
    ref = oracle.convolution_resample(kin, kout, v.astype(np.float32).astype(np.float64), kernel="b13")
    got32 = ci.convolution_resample(kin, kout, v.astype(np.float32), kernel="b13")
    assert got32.dtype == np.float32
    assert rel_err(got32, ref, 1.0).max() <= 1e-5
    got32 = ci.convolution_smooth(ks, v2.astype(np.float32), 0.05)
    assert rel_err(got32, oracle.convolution_smooth(ks, v2.astype(np.float32).astype(np.float64), 0.05), 1.0).max() <= 1e-5


def test_output_slab_decomposition_is_bit_identical(ci):
    """Multi-GPU formulation of the sweeps: replicated input, output split into slabs along the last
    axis.  Emulated here rank by rank on one device: the concatenated slabs equal the single-device result bitwise."""
    rng = np.random.default_rng(25)
    kin = tuple(np.linspace(0, 1, n) for n in (30, 28, 33))
    kout = tuple(np.linspace(0, 1, n) for n in (40, 50, 61))
    v = rng.random((30, 28, 33))
    full = ci.convolution_resample(kin, kout, v, kernel="b13", derivative=(1, 0, 0))
    parts = []
    for r in range(4):
        ks, _ = ci.sharding.output_slab(kout, 4, r)
        parts.append(ci.convolution_resample(kin, ks, v, kernel="b13", derivative=(1, 0, 0)))
    assert np.array_equal(np.concatenate(parts, axis=2), full)


def test_resample_block_table_cache_and_tensor_core_consumer_shapes(ci, oracle):
    """The dense block tables of a pass are cached across calls, keyed on the output knots and the kernel
    (cib200_resample.cu, block_tables_cached): a repeated call returns the same bits, a call onto other knots (same
    lengths, shifted values) must not see the first call's tables, and more distinct knot sets than the cache holds
    (48 entries, LRU eviction) still match the oracle.  Shapes: tiles with fewer than 32 valid lines, output lengths that
    are not a multiple of the 8-knot blocks, a window bound that changes with the knots (tensor-core consumer, K steps)."""
    rng = np.random.default_rng(91)
    kin = (np.linspace(0.0, 1.0, 37), np.linspace(-1.0, 2.0, 50), np.linspace(0.0, 3.0, 44))
    v = _sin_field((37, 50, 44), rng, 0.02)
    kout = (np.linspace(0.0, 1.0, 61), np.linspace(-1.0, 2.0, 83), np.linspace(0.0, 3.0, 70))
    a, _ = _check_resample(ci, oracle, kin, kout, v, kernel="b5")
    b = ci.convolution_resample(kin, kout, v, kernel="b5")
    assert np.array_equal(a, b)
    kout2 = (np.linspace(0.02, 0.97, 61), np.linspace(-0.9, 1.8, 83), np.linspace(0.1, 2.9, 70))
    c, _ = _check_resample(ci, oracle, kin, kout2, v, kernel="b5")
    assert not np.array_equal(a, c)
    _check_resample(ci, oracle, kin, kout, v, kernel="b5", derivative=(0, 1, 0))  # another table pair, same knots
    x = np.linspace(0.0, 1.0, 40)
    v2 = _sin_field((40, 36), rng, 0.02)
    for i in range(60):  # more knot sets than cache entries
        ko = (np.linspace(0.001 * i, 1.0 - 0.0007 * i, 57 + (i % 3)), np.linspace(0.0, 1.0, 36)[: 36] * 1.0)
        got = ci.convolution_resample((x, np.linspace(0.0, 1.0, 36)), ko, v2, kernel="b7")
        if i % 10 == 0 or i == 59:
            ref = oracle.convolution_resample((x, np.linspace(0.0, 1.0, 36)), ko, v2, kernel="b7")
            assert rel_err(got, ref, np.abs(v2).max()).max() <= TOL64
    again = ci.convolution_resample(kin, kout, v, kernel="b5")  # evicted in between: rebuilt, same bits
    assert np.array_equal(a, again)
