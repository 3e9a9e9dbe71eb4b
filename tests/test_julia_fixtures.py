"""The Julia parity recipe (tools/julia_cases.py + tools/julia_dump.jl) end to end against a stub producer, and -
when somebody has run the real reference and committed its outputs under tests/golden/julia/ - the actual pin."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "julia")


def test_fixture_pipeline_runs_end_to_end_against_the_stub(tmp_path, oracle):
    from tools import julia_cases
    root = str(tmp_path / "cases")
    julia_cases.write_cases(root, small=True)
    names = sorted(os.listdir(root))
    assert {"c1_detect", "c1_poly", "c2_line", "c3_d000", "c3_d100", "c4_smooth", "c4_smooth_plus5", "smooth_quirk_cos",
            "smooth_quirk_cos_plus5", "c5_poly", "c5_detect", "b13_d100_flat"} <= set(names)
    # nothing to check before a producer ran
    assert julia_cases.check_cases(root, {"oracle": oracle}) == {}
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "julia_stub.py"), root], check=True)
    rep = julia_cases.check_cases(root, {"oracle": oracle})
    assert set(rep) == set(n for n in names if os.path.isdir(os.path.join(root, n)))
    assert all(v == 0.0 for r in rep.values() for v in r.values())
    # a corrupted producer output is caught
    p = os.path.join(root, "c5_poly", "out.f64")
    out = np.fromfile(p, dtype="<f8")
    out[17] += 1e-9 * np.abs(np.fromfile(os.path.join(root, "c5_poly", "coefs.f64"), dtype="<f8")).max()
    out.tofile(p)
    with pytest.raises(AssertionError):
        julia_cases.check_cases(root, {"oracle": oracle})


def test_julia_dump_script_reads_what_the_case_writer_writes():
    """julia_dump.jl cannot run here; its file names and spec keys are checked against the writer's."""
    src = open(os.path.join(ROOT, "tools", "julia_dump.jl")).read()
    for token in ('"spec.txt"', '"values.f64"', '"knots_$(d-1).f64"', '"x_$(d-1).f64"', '"knots_out_$(d-1).f64"',
                  '"out.f64"', '"coefs.f64"', 'sp["kind"]', 'sp["kernel"]', 'sp["derivative"]', 'sp["bc"]',
                  'sp["extrap"]', 'sp["subgrid"]', 'sp["B"]', 'sp["ndim"]', 'sp["n"]'):
        assert token in src, token


@pytest.mark.skipif(not os.path.isdir(GOLDEN), reason="no reference outputs committed (Julia has never been available)")
def test_oracle_against_real_julia_outputs(oracle):
    from tools import julia_cases
    rep = julia_cases.check_cases(GOLDEN, {"oracle": oracle})
    assert rep
    stamp = open(os.path.join(GOLDEN, "JULIA_VERSION")).read()
    assert "STUB" not in stamp


@pytest.mark.gpu
def test_cuda_library_runs_the_fixture_cases(tmp_path, oracle, ci):
    """the same cases through the C-ABI (and against real Julia outputs when tests/golden/julia exists)"""
    from tools import julia_cases
    root = str(tmp_path / "cases")
    julia_cases.write_cases(root, small=True)
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "julia_stub.py"), root], check=True)
    rep = julia_cases.check_cases(root, {"cuda": ci})
    assert rep and all(v <= 1e-12 for r in rep.values() for v in r.values())
    if os.path.isdir(GOLDEN):
        julia_cases.check_cases(GOLDEN, {"cuda": ci})
