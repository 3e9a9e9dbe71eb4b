#!/usr/bin/env python3
"""AUTO's density thresholds (measured on 512^3) checked on smaller grids: whole call in ms, AUTO against the forced
alternatives (experiments build: CIB200_SO_NAME=libcib200_exp.so)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

rng = np.random.default_rng(2)
for n in (96, 128, 256):
    k = tuple(np.linspace(0, 1, n) for _ in range(3))
    v = rng.random((n, n, n))
    itp = ci.convolution_interpolation(k, v, kernel="b5", bc="poly", eval_mode="binned")
    itp.itp.set_profiling(True)
    ncell = (n - 1) ** 3
    for dens in (0.5, 1.0, 2.0, 4.0, 8.0, 16.0):
        nq = int(dens * ncell)
        q = [torch.rand(nq, dtype=torch.float64, device="cuda") for _ in range(3)]
        out = torch.empty(nq, dtype=torch.float64, device="cuda")
        row = []
        for env in ({}, {"CIB200_LANEQUERY": "0"}, {"CIB200_LANEQUERY": "3"}, {"CIB200_LANEQUERY": "4"}, {"CIB200_LANEQUERY": "0", "CIB200_BINNED_CELLSORT": "0"}):
            for kk in ("CIB200_LANEQUERY", "CIB200_BINNED_CELLSORT"):
                os.environ.pop(kk, None)
            os.environ.update(env)
            best = 1e30
            for _ in range(4):
                itp.eval(*q, out=out)
                torch.cuda.synchronize()
                best = min(best, sum(itp.itp.last_pass_ms().values()))
            row.append((best, itp.itp.last_strategy()))
        print(f"{n}^3 {dens:5.1f}/cell nq {nq:9d}: AUTO {row[0][1]:14s} {row[0][0]:7.3f} | cellrun {row[1][0]:7.3f} | pairs {row[2][0]:7.3f} | blocks {row[3][0]:7.3f} | binned {row[4][0]:7.3f}", flush=True)
    itp.itp.close()
