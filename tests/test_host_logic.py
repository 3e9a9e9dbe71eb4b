"""CPU-side checks of the product: the C-ABI library loads and exports every declared symbol, its embedded
tables are the reference's, the HOST-side coefficient builder (north_star subsystem 1) reproduces the oracle
bit for bit, and the host mirror resolves modes / rejects unsupported ones like the reference.  No compute
call needs a GPU here."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(ci):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "cib200.h")).read()
    declared = set(re.findall(r"^(?:int32_t|int64_t|const char \*|const double \*)\s*(cib200_[a-z0-9_]+)\(", header, re.M))
    assert declared == set(ci.SYMBOLS), declared ^ set(ci.SYMBOLS)
    L = ci.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert L.cib200_version() >= 100


def test_embedded_tables_are_the_shipped_ones(ci, golden_dir):
    g = np.load(os.path.join(golden_dir, "shipped_tables.npz"))
    ids = {"b5": 0, "b7": 1, "b9": 2, "b11": 3, "b13": 4, "a3": 5, "a4": 6, "a5": 7, "a7": 8}
    L = ci.lib()
    n = 0
    for name in g.files:
        m = re.match(r"KERNEL_(D1_|D2_|)PRE_([ab]\d+)_(d|i)(\d+)$", name)
        base = int(m.group(4)) if m.group(3) == "d" else -int(m.group(4))  # _i1: antiderivative tables (deriv = -1)
        kernel, d = m.group(2), base + {"": 0, "D1_": 1, "D2_": 2}[m.group(1)]
        eqs = C.c_int32()
        p = L.cib200_kernel_table(ids[kernel], d, C.byref(eqs))
        t = np.ctypeslib.as_array(p, shape=(101, 2 * eqs.value))
        lit = g[name]
        assert ((t.view(np.uint64) == lit.view(np.uint64)) | ((t == 0) & (lit == 0))).all(), name
        n += 1
    assert n == 102
    assert not L.cib200_kernel_table(0, 4, None)  # b5 has no 4th derivative table


def _build(ci, values, h, kernels, bcs, dtype=np.float64):
    from convitp_b200 import _lib
    v = np.asfortranarray(values, dtype=dtype)
    nd = v.ndim
    eqs = [{"b5": 5, "b7": 6, "b9": 7, "b11": 8, "b13": 9, "a3": 2, "a4": 3, "a5": 3, "a7": 4}[k] for k in kernels]
    out = np.zeros(tuple(s + 2 * (e - 1) for s, e in zip(v.shape, eqs)), dtype=dtype, order="F")
    st = ci.lib().cib200_build_coefs(
        nd, 0 if dtype == np.float64 else 1, (C.c_int64 * nd)(*v.shape), v.ctypes.data, (C.c_double * nd)(*h),
        (C.c_int32 * nd)(*[_lib.KERNEL_IDS[k] for k in kernels]),
        (C.c_int32 * nd)(*[_lib.BC_IDS[b[0]] for b in bcs]), (C.c_int32 * nd)(*[_lib.BC_IDS[b[1]] for b in bcs]),
        out.ctypes.data)
    _lib.check(st)
    return out


@pytest.mark.parametrize("shape,kernels", [((40,), ("b5",)), ((7,), ("b13",)), ((17, 23), ("b5", "b5")),
                                          ((12, 9), ("b13", "b5")), ((14, 11, 13), ("b5", "b5", "b5")),
                                          ((10, 12, 9), ("b13", "b7", "b9")),
                                          ((9,), ("a3",)), ((15, 12), ("a4", "a7")), ((8, 9, 10), ("a5", "b5", "a4"))])
@pytest.mark.parametrize("bc", ["detect", "poly", "linear", "quadratic", ("detect", "linear")])
def test_host_coefficient_builder_bit_equals_oracle(ci, oracle, shape, kernels, bc):
    rng = np.random.default_rng(hash((shape, kernels, str(bc))) % 2 ** 32)
    # smooth field + a little noise so that :detect takes both branches across lines
    grids = np.meshgrid(*[np.linspace(0, 3, n) for n in shape], indexing="ij")
    vals = sum(np.sin(1.3 * g + i) for i, g in enumerate(grids)) + 0.02 * rng.standard_normal(shape)
    h = [3.0 / (n - 1) for n in shape]
    bcs = tuple((bc, bc) if isinstance(bc, str) else tuple(bc) for _ in shape)
    got = _build(ci, vals, h, kernels, bcs)
    ref = oracle.create_coefs(vals, h, kernels, bcs)
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


def test_detect_takes_both_branches(oracle):
    # sanity of the fixture above: :detect must choose polynomial on smooth lines and linear on noisy ones
    x = np.linspace(0, 3, 40)
    smooth = oracle.create_coefs(np.sin(x), [x[1] - x[0]], ("b5",), (("detect", "detect"),))
    poly = oracle.create_coefs(np.sin(x), [x[1] - x[0]], ("b5",), (("poly", "poly"),))
    lin = oracle.create_coefs(np.sin(x), [x[1] - x[0]], ("b5",), (("linear", "linear"),))
    assert np.array_equal(smooth, poly) and not np.array_equal(smooth, lin)
    rough = np.sin(x) + 0.3 * np.random.default_rng(0).standard_normal(40)
    assert np.array_equal(oracle.create_coefs(rough, [x[1] - x[0]], ("b5",), (("detect", "detect"),)),
                          oracle.create_coefs(rough, [x[1] - x[0]], ("b5",), (("linear", "linear"),)))


def test_extend_knots_matches_extend_vector(ci, oracle):
    k = np.linspace(-1.0, 2.5, 33)
    out = np.zeros(33 + 16)
    from convitp_b200 import _lib
    _lib.check(ci.lib().cib200_extend_knots(k.ctypes.data, 33, 8, out.ctypes.data))
    assert np.array_equal(out, oracle.extend_vector(k, 8))


def test_f32_builder_rounds_the_f64_result(ci, oracle):
    rng = np.random.default_rng(5)
    v = rng.random((20, 21)).astype(np.float32)
    got = _build(ci, v, [0.1, 0.1], ("b5", "b5"), (("poly", "poly"),) * 2, dtype=np.float32)
    ref = oracle.create_coefs(v.astype(np.float64), [0.1, 0.1], ("b5", "b5"), (("poly", "poly"),) * 2)
    # axis-1 ghosts are computed from float-rounded axis-0 ghosts: compare with a float tolerance
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max()
    assert np.array_equal(got[4:-4, 4:-4], v)


def test_unsupported_modes_are_rejected_not_approximated(ci):
    x = np.linspace(0, 1, 30)
    v = np.sin(x)
    U = ci.UnsupportedModeError
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, kernel="a1")           # a0 / a1: :linear subgrid only
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, kernel="a4", derivative=2)  # K'' of the a-series does not exist
    with pytest.raises(U):  # mixed integral / derivative axes: K^(4) of :b5 does not exist
        ci.convolution_interpolation((x, x), np.zeros((30, 30)), derivative=(-1, 4))
    with pytest.raises(U):  # the antiderivative path is Float64 only
        ci.convolution_interpolation(x, v, derivative=-1, dtype=np.float32)
    with pytest.raises(U):  # N >= 4 integral directions: :linear subgrid
        ci.convolution_interpolation((x[:12],) * 4, np.zeros((12,) * 4), derivative=(-1, 0, 0, -1))  # < 3 integral axes in 4-D
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, derivative=4)          # K^(4) of :b5 does not exist in the tables
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, derivative=3, precompute=1000)  # generated tables: reference path
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, derivative=2, subgrid="quintic")   # K^(4) of :b5 is not shipped
    with pytest.raises(U):
        v2 = np.zeros((30, 30))
        ci.convolution_interpolation((x, x), v2, subgrid=("quintic", "cubic"))  # eager 2-D: (quintic, quintic) only
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, lazy=True)
    with pytest.raises(U):
        ci.convolution_interpolation(x, v, fast=False)
    with pytest.raises(U):
        ci.convolution_interpolation(np.sort(np.random.default_rng(0).random(30)), v)  # non-uniform grid
    with pytest.raises(U):
        ci.convolution_interpolation((x[:4],) * 9, np.zeros((4,) * 9), kernel="a3")  # N > CIB200_MAX_ND = 8
    with pytest.raises(U):
        ci.convolution_interpolation((x[:8],) * 4, np.zeros((8,) * 4), kernel="a4", derivative=2)  # 4-D: K'' of a4
    with pytest.raises(U):  # generated tables
        ci.convolution_resample(x, x, v, kernel="b5", precompute=1000)
    with pytest.raises(ValueError, match="antiderivatives"):
        ci.convolution_resample(x, x, v, derivative=-1)
    with pytest.raises(ValueError, match="Cannot use cubic subgrid with a1"):
        ci.convolution_resample(x, x, v, kernel="a1", subgrid="cubic")
    with pytest.raises(ValueError, match="half-width"):
        ci.convolution_smooth(np.linspace(0, 1, 20), np.zeros(20), 0.05)
    with pytest.raises(ValueError, match="Unknown extrapolation"):
        ci.convolution_interpolation(x, v, extrap="bogus")


def test_mode_resolution_mirrors_reference(ci):
    from convitp_b200 import api
    assert api._build_subgrids(("b5", "b5"), (3, 0), ("cubic", "cubic")) == ("linear", "cubic")
    assert api._effective_subgrids(("linear", "cubic", "cubic"), 3) == ("linear",) * 3  # perdim_3d.jl:161
    assert api._effective_subgrids(("linear", "cubic"), 2) == ("linear", "cubic")
    assert api._build_subgrids(("b5", "b5"), (3, 3), ("cubic", "cubic")) == ("linear", "linear")
    assert api._build_subgrids(("b13",) * 3, (1, 0, 0), ("cubic",) * 3) == ("cubic",) * 3
    assert api._build_subgrids(("b5",) * 4, (0,) * 4, ("cubic",) * 4) == ("linear",) * 4
    assert api._norm_bc("poly", 2) == (("poly", "poly"), ("poly", "poly"))
    assert api._norm_bc(("detect", "linear"), 3) == (("detect", "linear"),) * 3
    assert api._extrap_type(":line") == api.Line() and api._extrap_type(api.Flat()) == api.Flat()
    assert api.is_uniform_grid(np.linspace(0, 1, 50)) and not api.is_uniform_grid(np.array([0, 1, 2.1, 3]))


def test_no_cpu_fallback_without_a_device(ci):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    x = np.linspace(0, 1, 30)
    with pytest.raises(ci.CIB200Error) as e:
        ci.convolution_interpolation(x, np.sin(x))
    assert e.value.status == 5  # CIB200_ECUDA
    with pytest.raises(ci.CIB200Error) as e:
        ci.convolution_interpolation(x, np.sin(x), derivative=-1)  # antiderivative handle (cib200_itg_create)
    assert e.value.status == 5
    with pytest.raises(ci.CIB200Error):
        ci.convolution_smooth(np.linspace(0, 1, 100), np.zeros(100), 0.05)
    with pytest.raises(ci.CIB200Error):
        ci.convolution_resample(x, x, np.sin(x))


def test_linear_subgrid_axes_warn_about_the_table_length(ci):
    """The reference raises :linear axes to >= 10 000 table rows (constructor.jl:117-118); this library evaluates them on
    the 101-row shipped tables and must say so (DESIGN.md 1, row f) - checked on the mode resolution, no device needed."""
    from convitp_b200 import api
    with pytest.warns(ci.LinearSubgridTableWarning):
        api._warn_linear_rows(("linear", "cubic"), ("b5", "b5"), "test")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        api._warn_linear_rows(("cubic", "quintic"), ("b5", "b5"), "test")
        api._warn_linear_rows(("linear",), ("a1",), "test")  # a0 / a1 keep their two-row tables


def test_generated_tables_of_linear_axes(ci, oracle, golden_dir):
    """tables.generated_table = precompute_kernel (precompute_kernel_and_range.jl:150-157) in exact integer arithmetic: 101
    rows reproduce the shipped literals bit for bit, 10 000 rows equal the oracle's independent restatement, and a row of
    the long table that coincides with a shipped sample point carries the same bits"""
    from convitp_b200 import tables as T
    g = np.load(os.path.join(golden_dir, "shipped_tables.npz"))
    for k, d in (("b5", 0), ("b5", 3), ("b9", 2), ("b13", 6), ("a4", 0), ("a7", 1)):
        name = f"KERNEL_PRE_{k}_d{d}" if f"KERNEL_PRE_{k}_d{d}" in g.files else f"KERNEL_D1_PRE_{k}_d{d - 1}"
        t = T.generated_table(k, d, 101)
        lit = g[name]
        assert ((t.view(np.uint64) == lit.view(np.uint64)) | ((t == 0) & (lit == 0))).all(), (k, d)
    for k, d in (("b5", 3), ("a4", 1)):
        t = T.generated_table(k, d)
        assert t.shape == (10_000, 2 * T._polys()[k]["eqs"])
        assert np.array_equal(t, oracle.generated_table(k, d))
        short = T.generated_table(k, d, 101)
        assert np.array_equal(t[0], short[0]) and np.array_equal(t[-1], short[-1])      # s = c - eqs, c + 1 - eqs
        assert np.array_equal(t[1111 * 9], T.generated_table(k, d, 12)[1 * 11])  # 9999/9999 ... (r-1)/(n-1) = 1: last row again
