#!/usr/bin/env python3
"""Gather vs binned strategy by batch size: 3-D :b5, device-resident uniformly random queries, grids 512^3 and 128^3."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402


def t_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    return min(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))


for n in (512, 128):
    g = np.linspace(0, 1, n)
    v = np.random.default_rng(0).random((n, n, n))
    itps = {m: ci.convolution_interpolation((g, g, g), v, kernel="b5", bc="poly", eval_mode=m, coefs_on="device") for m in ("gather", "binned")}
    for lg in range(14, 25, 2):
        nq = 1 << lg
        xs = [torch.rand(nq, dtype=torch.float64, device="cuda") for _ in range(3)]
        out = torch.empty(nq, dtype=torch.float64, device="cuda")
        r = {m: t_ms(lambda: itps[m].eval(*xs, out=out)) for m in itps}
        print(f"grid {n}^3 nq 2^{lg}: gather {r['gather']:.3f} ms  binned {r['binned']:.3f} ms  ratio {r['gather'] / r['binned']:.2f}", flush=True)
