#!/usr/bin/env python3
"""ncu driver for the headline launch itself: 3-D :b5 on the 512^3 Float64 grid of bench.py, N uniformly random
device-resident queries (default 10^9: tile binning + cell sort + eval_cellrun_kernel<5>), after one warm-up call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import convitp_b200 as ci  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
kernel = sys.argv[2] if len(sys.argv) > 2 else "b5"
knots, vals = bench.make_grid(512)
itp = ci.convolution_interpolation(knots, vals, kernel=kernel, bc="poly")
del vals
gen = torch.Generator(device="cuda")
gen.manual_seed(bench.SEED)
xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
out = torch.empty(nq, dtype=torch.float64, device="cuda")
for _ in range(2):
    itp.eval(*xs, out=out)
torch.cuda.synchronize()
print("ok", float(out[:1000].sum()))
