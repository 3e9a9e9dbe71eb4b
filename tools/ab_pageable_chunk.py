#!/usr/bin/env python3
"""Pageable host buffers (plain numpy arrays): chunk size of the staged pipeline, experiments build."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, convitp_b200 as ci
nq = 500_000_000
knots, vals = bench.make_grid(512)
itp = ci.convolution_interpolation(knots, vals, kernel="b5", bc="poly")
del vals
rng = np.random.default_rng(1)
px = [rng.random(nq) for _ in range(3)]
pout = np.empty(nq)
for lg in (24, 25, 26):
    os.environ["CIB200_HOST_CHUNK_LOG2"] = str(lg)
    itp.eval(*px, out=pout)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter(); itp.eval(*px, out=pout); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    print(f"pageable, chunk 2^{lg}: {nq / best:.4e} evals/s ({best*1e3:.1f} ms), last chunk {itp.itp.last_strategy()}", flush=True)
