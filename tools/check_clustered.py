#!/usr/bin/env python3
"""Adversarial query distributions for the binned strategies: everything in one cell, one tile, one plane, a
Gaussian blob - results must equal the gather strategy (<= 1e-13 relative to max|v|), timings reported."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

n = 256
nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
rng = np.random.default_rng(7)
k = tuple(np.linspace(0, 1, n) for _ in range(3))
v = rng.random((n, n, n))
gb = ci.convolution_interpolation(k, v, bc="poly", eval_mode="binned")
gg = ci.convolution_interpolation(k, v, bc="poly", eval_mode="gather")
dev = torch.device("cuda")
h = 1.0 / (n - 1)


def cases():
    u = lambda: torch.rand(nq, dtype=torch.float64, device=dev)
    yield "uniform", [u(), u(), u()]
    yield "one cell", [0.5 + h * u(), 0.25 + h * u(), 0.75 + h * u()]
    yield "one tile", [0.5 + 16 * h * u(), 0.25 + 16 * h * u(), 0.75 + 16 * h * u()]
    yield "one z-plane of cells", [u(), u(), 0.5 + h * u()]
    g = lambda: (0.5 + 0.02 * torch.randn(nq, dtype=torch.float64, device=dev)).clamp_(0, 1)
    yield "gaussian blob sigma=0.02", [g(), g(), g()]


for name, q in cases():
    out = gb.eval(*q)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = gb.eval(*q)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sub = slice(0, nq, max(1, nq // 2_000_000))
    ref = gg.eval(*[x[sub].contiguous() for x in q])
    err = float((out[sub] - ref).abs().max())
    print(f"{name:28s} {nq / dt / 1e9:6.2f} e9 evals/s   max |binned - gather| = {err:.2e}")
    assert err <= 1e-13
