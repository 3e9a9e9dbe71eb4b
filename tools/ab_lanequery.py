#!/usr/bin/env python3
"""A/B of the cell-sorted evaluation kernels on the headline workload (3-D :b5, 512^3 Float64 grid, N uniformly random
device-resident queries): eval_cellrun_kernel against the lane-per-query kernel and its variants.  Needs an experiments
build (CIB200_SO_NAME=libcib200_exp.so python build.py --experiments): the switches are read from the environment there.

usage: CIB200_SO_NAME=libcib200_exp.so python tools/ab_lanequery.py [nq] [grid]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import convitp_b200 as ci  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
knots, vals = bench.make_grid(n)
itp = ci.convolution_interpolation(knots, vals, kernel="b5", bc="poly", eval_mode="binned")
vmax = float(abs(vals).max())
del vals
gen = torch.Generator(device="cuda")
gen.manual_seed(bench.SEED)
xs = [torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
out = torch.empty(nq, dtype=torch.float64, device="cuda")
itp.itp.set_profiling(True)
ref = None
cases = [("cellrun", "0", "0"), ("lq2 256thr", "2", "0"), ("mma 384thr breg", "3", "0"), ("mma 384thr", "3", "1"), ("mma 512thr", "3", "2"), ("mma 512thr breg", "3", "3"), ("mma blocks", "4", "0")]
only = os.environ.get("AB_CASES")  # comma-separated case names
reps = int(os.environ.get("AB_REPS", "3"))
for name, lq, var in cases:
    if only and name not in only.split(","):
        continue
    os.environ["CIB200_LANEQUERY"] = lq
    os.environ["CIB200_LQ_VARIANT"] = var
    os.environ["CIB200_MMA_VARIANT"] = var
    best = None
    for _ in range(reps):
        itp.eval(*xs, out=out)
        torch.cuda.synchronize()
        ms = itp.itp.last_pass_ms()
        if best is None or ms["eval_kernel"] < best["eval_kernel"]:
            best = ms
    if ref is None:
        ref = out.clone()
        diff = 0.0
    else:
        diff = float((out - ref).abs().max()) / vmax
    print(f"{name:20s} binning {best["binning"]:7.2f}  cellsort {best["cellsort"]:7.2f}  eval {best["eval_kernel"]:7.2f}  spill {best["spill"]:6.2f} ms   "
          f"max|d|/max|v| vs cellrun {diff:.2e}", flush=True)

if os.environ.get("AB_SPARSE"):
    # sparse batches: eval_binned_kernel (no cell sort) against cell sort + lane-per-query kernel, one query per lane
    print("sparse batches on the same grid: whole call (binning + cell sort + evaluation), ms")
    for lg in (18, 20, 22, 24, 26, 27, 28, 29):
        m = 1 << lg
        if m > nq:
            break
        q = [x[:m] for x in xs]
        row = []
        ref_s = None
        for cs, lq in (("0", "0"), ("1", "0"), ("1", "4"), ("1", "3")):
            os.environ["CIB200_BINNED_CELLSORT"] = cs
            os.environ["CIB200_LANEQUERY"] = lq
            os.environ["CIB200_LQ_VARIANT"] = "0"
            os.environ["CIB200_MMA_VARIANT"] = "0"
            best = 1e30
            for _ in range(3):
                itp.eval(*q, out=out[:m])
                torch.cuda.synchronize()
                ms = itp.itp.last_pass_ms()
                best = min(best, sum(ms.values()))
            if ref_s is None:
                ref_s = out[:m].clone()
            else:
                row_err = float((out[:m] - ref_s).abs().max()) / vmax
                assert row_err < 1e-13, (lg, cs, lq, row_err)
            row.append(best)
        print(f"2^{lg}: eval_binned {row[0]:8.3f}   cellsort+cellrun {row[1]:8.3f}   cellsort+mma(blocks) {row[2]:8.3f}   cellsort+mma(pairs) {row[3]:8.3f}", flush=True)
