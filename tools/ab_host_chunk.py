#!/usr/bin/env python3
"""Host-buffer (e2e) leg of bench.py for several chunk sizes of the H2D -> kernels -> D2H pipeline, experiments build:
CIB200_SO_NAME=libcib200_exp.so python tools/ab_host_chunk.py [nq]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import convitp_b200 as ci  # noqa: E402

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
knots, vals = bench.make_grid(512)
itp = ci.convolution_interpolation(knots, vals, kernel="b5", bc="poly")
del vals
gen = torch.Generator(device="cuda")
gen.manual_seed(bench.SEED)
hx = [torch.empty(nq, dtype=torch.float64, pin_memory=True) for _ in range(3)]
for i in range(3):
    hx[i].copy_(torch.rand(nq, dtype=torch.float64, device="cuda", generator=gen))
hout = torch.empty(nq, dtype=torch.float64, pin_memory=True)
hxn, houtn = [h.numpy() for h in hx], hout.numpy()
for lg in (26, 27, 28, 29):
    os.environ["CIB200_HOST_CHUNK_LOG2"] = str(lg)
    itp.eval(*hxn, out=houtn)
    best = 1e30
    for _ in range(2):
        t0 = time.perf_counter()
        itp.eval(*hxn, out=houtn)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"chunk 2^{lg}: {nq / best:.4e} evals/s  ({best * 1e3:.1f} ms per 10^9-query step, last chunk strategy {itp.itp.last_strategy()})", flush=True)
