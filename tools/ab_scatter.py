#!/usr/bin/env python3
"""A/B of the scatter pass of the tile binning (private / shared write cursors, L2 eviction hints) on the headline workload,
experiments build: CIB200_SO_NAME=libcib200_exp.so python tools/ab_scatter.py [nq]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import convitp_b200 as ci  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
knots, vals = bench.make_grid(512)
itp = ci.convolution_interpolation(knots, vals, kernel="b5", bc="poly", eval_mode="binned")
del vals
gen = torch.Generator(device="cuda")
gen.manual_seed(bench.SEED)
xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
out = torch.empty(nq, dtype=torch.float64, device="cuda")
itp.itp.set_profiling(True)
ref = None
for env in ({}, {"CIB200_SCATTER_SHARED": "1"}, {"CIB200_SCATTER_HINTS": "1"}, {"CIB200_SCATTER_SHARED": "1", "CIB200_SCATTER_HINTS": "1"},
            {"CIB200_CELLSORT_L2HINT": "1"}, {"CIB200_CELLSORT_PERM": "0"}, {}):
    for k in ("CIB200_SCATTER_SHARED", "CIB200_SCATTER_HINTS", "CIB200_CELLSORT_L2HINT", "CIB200_CELLSORT_PERM"):
        os.environ.pop(k, None)
    os.environ.update(env)
    best = None
    for _ in range(3):
        itp.eval(*xs, out=out)
        torch.cuda.synchronize()
        ms = itp.itp.last_pass_ms()
        if best is None or sum(ms.values()) < sum(best.values()):
            best = ms
    if ref is None:
        ref = out.clone()
    same = bool(torch.equal(ref, out))
    print(f"{str(env):70s} binning {best['binning']:6.2f}  cellsort {best['cellsort']:6.2f}  eval {best['eval_kernel']:7.2f}  total {sum(best.values()):7.2f} ms  same bits {same}", flush=True)
