"""The CPU oracle satisfies every analytic property the reference's own test-suite asserts for the hot
path.  No GPU needed."""
import types

import numpy as np
import pytest

from . import reference_properties as RP


@pytest.fixture(scope="module")
def impl(oracle):
    return types.SimpleNamespace(interp=oracle.convolution_interpolation, resample=oracle.convolution_resample,
                                 smooth=oracle.convolution_smooth)


@pytest.mark.parametrize("check", RP.ALL_CHECKS, ids=lambda f: f.__name__)
def test_reference_property(check, impl):
    check(impl)


def test_bc_downgrade_and_subgrid_resolution(oracle):
    # test_boundary_condition.jl:3-31
    xs = np.linspace(0, 2 * np.pi, 5)
    assert oracle.Interpolant(xs, np.sin(xs), kernel="b5", bc="poly").bc == (("linear", "linear"),)
    v3 = np.sin(xs)[:, None, None] * np.cos(xs)[None, :, None] * np.sin(xs)[None, None, :]
    assert oracle.Interpolant((xs,) * 3, v3, kernel="b5", bc="poly").bc == (("linear", "linear"),) * 3
    # test_subgrid_downgrade.jl:2-17
    assert oracle.build_subgrids(("b5", "b5"), (3, 0), ("cubic", "cubic")) == ("linear", "cubic")
    assert oracle.build_subgrids(("b5", "b5"), (3, 3), ("cubic", "cubic")) == ("linear", "linear")
    assert oracle.build_subgrids(("b5",) * 3, (1, 0, 0), ("cubic",) * 3) == ("cubic",) * 3


def test_throw_and_nan(oracle):
    xs = np.linspace(0, 1, 20)
    itp = oracle.Interpolant(xs, xs ** 2)
    with pytest.raises(oracle.OracleError, match="outside"):
        itp.eval(np.array([0.5, 1.5]))
    with pytest.raises(oracle.OracleError, match="InexactError"):
        itp.eval(np.array([0.5, np.nan]))
    assert itp(1.0) == pytest.approx(1.0, abs=1e-12)  # upper bound inclusive (extrapolation_operator.jl:43)


def test_smooth_boundary_quirks(oracle):
    """ghosts are the CENTRED least-squares line shifted one slot outward, so smoothing is not
    invariant under adding a constant (cos: 0.120 at the last point; cos+5: 2.30)."""
    xs = np.linspace(0, 2 * np.pi, 200)
    e0 = np.abs(oracle.convolution_smooth(xs, np.cos(xs), 0.05) - np.cos(xs))
    e5 = np.abs(oracle.convolution_smooth(xs, np.cos(xs) + 5, 0.05) - np.cos(xs) - 5)
    assert 0.11 < e0.max() < 0.13 and e0.argmax() in (0, 199)
    assert 2.2 < e5.max() < 2.4
    assert oracle.smooth_halfwidth(0.05) == 24 and oracle.smooth_halfwidth(0.02) == 38 and oracle.smooth_halfwidth(0.1) == 17
    with pytest.raises(oracle.OracleError, match="half-width"):
        oracle.convolution_smooth(np.linspace(0, 1, 20), np.zeros(20), 0.05)
