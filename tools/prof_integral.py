"""Antiderivative path (cib200_itg_*): construction (prefix-sum arrays on the device) and evaluation timings, also the
command profiled by ncu.  Usage: python tools/prof_integral.py [1d|2d|3d|all] [n3d] [nq3d]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n3 = int(sys.argv[2]) if len(sys.argv) > 2 else 256
nq3 = int(float(sys.argv[3])) if len(sys.argv) > 3 else 20_000_000


def timed(itp, qs, reps=5):
    for _ in range(2):
        itp.eval(*qs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        itp.eval(*qs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def build(knots, v, **kw):
    ci.convolution_interpolation(knots, v, derivative=-1, coefs_on="device", **kw)  # warm-up (allocator, module load)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    itp = ci.convolution_interpolation(knots, v, derivative=-1, coefs_on="device", **kw)
    torch.cuda.synchronize()
    return itp, (time.perf_counter() - t0) * 1e3


rng = np.random.default_rng(0)
if which in ("1d", "all"):
    x = np.linspace(0, 1, 1000)
    itp, bms = build(x, np.sin(2 * np.pi * x) + 0.1 * rng.uniform(-1, 1, 1000), kernel="b5")
    q = [torch.rand(1_000_000, dtype=torch.float64, device="cuda")]
    ms = timed(itp, q)
    print(f"1-D b5 n=1000 antiderivative: build {bms:.2f} ms (host wall), 1e6 queries {ms:.4f} ms = {1e6 / ms * 1e3:.3e} evals/s")
if which in ("2d", "all"):
    x = np.linspace(0, 2 * np.pi, 200)
    itp, bms = build((x, x), np.sin(x)[:, None] * np.cos(x)[None, :], kernel="b5")
    q = [torch.rand(10_000_000, dtype=torch.float64, device="cuda") * 2 * np.pi for _ in range(2)]
    ms = timed(itp, q)
    print(f"2-D b5 200^2 antiderivative: build {bms:.2f} ms, 1e7 queries {ms:.3f} ms = {1e7 / ms * 1e3:.3e} evals/s")
if which in ("3d", "all"):
    x = np.linspace(0, 1, n3)
    v = rng.random((n3, n3, n3))
    itp, bms = build((x, x, x), v, kernel="b5", bc="poly")
    q = [torch.rand(nq3, dtype=torch.float64, device="cuda") for _ in range(3)]
    ms = timed(itp, q, reps=3)
    print(f"3-D b5 {n3}^3 antiderivative: build {bms:.1f} ms (ghost fill + 7 prefix arrays of {(n3 + 8) ** 3 * 8 / 1e6:.0f} MB), "
          f"{nq3:.1e} queries {ms:.2f} ms = {nq3 / ms * 1e3:.3e} evals/s")
    # the plain interpolant on the same grid and queries, for scale
    it0 = ci.convolution_interpolation((x, x, x), v, kernel="b5", bc="poly", coefs_on="device")
    ms0 = timed(it0, q, reps=3)
    print(f"   same grid, derivative=0 (AUTO strategy): {ms0:.2f} ms = {nq3 / ms0 * 1e3:.3e} evals/s")
