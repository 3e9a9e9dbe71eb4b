#!/usr/bin/env python3
"""Context-independence check of the binned / cell-sorted evaluation: every query must get the same bits whatever
batch it is evaluated in.  Small enough to run under compute-sanitizer (racecheck / memcheck / initcheck)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
nq = int(float(sys.argv[2])) if len(sys.argv) > 2 else 400_000
rng = np.random.default_rng(5)
k = tuple(np.linspace(0, 1, n) for _ in range(3))
v = rng.random((n, n, n))
g = ci.convolution_interpolation(k, v, bc="poly", eval_mode="binned")
q = [torch.from_numpy(rng.random(nq)).cuda() for _ in range(3)]
full = g.eval(*q)
half = g.eval(*[x[: nq // 2].contiguous() for x in q])
third = g.eval(*[x[nq // 3:].contiguous() for x in q])
again = g.eval(*q)
gg = ci.convolution_interpolation(k, v, bc="poly", eval_mode="gather").eval(*q)
print("repeat equal:", bool(torch.equal(full, again)))
print("half equal:", bool(torch.equal(full[: nq // 2], half)), int((full[: nq // 2] != half).sum()))
print("third equal:", bool(torch.equal(full[nq // 3:], third)), int((full[nq // 3:] != third).sum()))
print("max |binned - gather|:", float((full - gg).abs().max()))
