#!/usr/bin/env python3
"""Stand-in for `julia tools/julia_dump.jl DIR` on boxes without Julia: writes out.f64 / coefs.f64 in the same
format from the CPU oracle.  It exists so that the fixture format, the loader and the comparison of
tools/julia_cases.py are exercised end to end by the tests; it pins nothing (the producer is the restatement
itself) and stamps the directory accordingly."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(root):
    from oracle import oracle as O
    from tools.julia_cases import read_case, run_case
    for name in sorted(os.listdir(root)):
        d = os.path.join(root, name)
        if not os.path.exists(os.path.join(d, "spec.txt")):
            continue
        out, _, coefs = run_case(O, read_case(d))
        np.asarray(out, dtype="<f8").tofile(os.path.join(d, "out.f64"))
        if coefs is not None:
            np.asarray(coefs, dtype="<f8").tofile(os.path.join(d, "coefs.f64"))
    open(os.path.join(root, "JULIA_VERSION"), "w").write("STUB (tools/julia_stub.py: the oracle, not Julia)")


if __name__ == "__main__":
    main(sys.argv[1])
