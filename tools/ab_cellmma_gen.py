#!/usr/bin/env python3
"""Timing of the generic tensor-core kernel (eval_cellmma_gen_kernel: any :b width, several table pairs) against the
strategies it replaces, on the 512^3 grid of bench.py.  Needs an experiments build for the forced-strategy columns
(CIB200_SO_NAME=libcib200_exp.so python build.py --experiments).

usage: CIB200_SO_NAME=libcib200_exp.so python tools/ab_cellmma_gen.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import convitp_b200 as ci  # noqa: E402

knots, vals = bench.make_grid(512)
gen = torch.Generator(device="cuda")
gen.manual_seed(bench.SEED)
NQ = 1_000_000_000
xs = [torch.rand(NQ, dtype=torch.float64, device="cuda", generator=gen) for _ in range(3)]
out = torch.empty(NQ, dtype=torch.float64, device="cuda")


def run(itp, nq, env):
    for k in ("CIB200_LANEQUERY", "CIB200_MMA_GENERIC", "CIB200_BINNED_CELLSORT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    best = None
    for _ in range(3):
        itp.eval(*[x[:nq] for x in xs], out=out[:nq])
        torch.cuda.synchronize()
        ms = itp.itp.last_pass_ms()
        tot = sum(ms.values())
        if best is None or tot < best[0]:
            best = (tot, ms["eval_kernel"], itp.itp.last_strategy())
    return best


for kw, nq in ((dict(kernel="b5", derivative=(1, 0, 0)), NQ), (dict(kernel="b5", derivative=(1, 0, 0)), 250_000_000), (dict(kernel="b7"), 500_000_000),
               (dict(kernel="b7"), 150_000_000), (dict(kernel="b9"), 250_000_000), (dict(kernel="b11"), 250_000_000),
               (dict(kernel="b13"), 125_000_000), (dict(kernel="b13"), 250_000_000), (dict(kernel="b13"), 500_000_000), (dict(kernel="b13"), 1_000_000_000)):
    itp = ci.convolution_interpolation(knots, vals, bc="poly", eval_mode="binned", **kw)
    itp.itp.set_profiling(True)
    auto = run(itp, nq, {})
    ref = out[:1 << 20].clone()
    alt = run(itp, nq, {"CIB200_LANEQUERY": "0"})
    d = float((out[:1 << 20] - ref).abs().max() / ref.abs().max())
    prs = run(itp, nq, {"CIB200_LANEQUERY": "3"})
    blk = run(itp, nq, {"CIB200_LANEQUERY": "4"})
    d2 = float((out[:1 << 20] - ref).abs().max() / ref.abs().max())
    line = (f"{kw} nq {nq:.1e}: default {auto[2]} {auto[0]:8.2f} ms (kernel {auto[1]:8.2f})   without the tensor-core kernel {alt[2]} {alt[0]:8.2f} ms "
            f"(kernel {alt[1]:8.2f})  max rel diff {d:.1e}   pairs {prs[0]:8.2f}   blocks {blk[2]} {blk[0]:8.2f} (diff {d2:.1e})")
    if kw == dict(kernel="b5"):
        g = run(itp, nq, {"CIB200_MMA_GENERIC": "1"})
        line += f"   generic kernel forced: {g[0]:8.2f} ms (kernel {g[1]:8.2f})"
    print(line, flush=True)
    itp.itp.close()
