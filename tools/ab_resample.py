#!/usr/bin/env python3
"""A/B of the convolution_resample pass kernels on BASELINE config 3 (:b13 128^3 -> 256^3), experiments build:
CIB200_SO_NAME=libcib200_exp.so python tools/ab_resample.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402
from tools.prof_sweeps import julia_order  # noqa: E402  (runs one warm-up of every sweep on import)

rng = np.random.default_rng(3)
kin = tuple(np.linspace(0, 2 * np.pi, 128) for _ in range(3))
kout = tuple(np.linspace(0, 2 * np.pi, 256) for _ in range(3))
v = julia_order(rng.random((128, 128, 128)))


def best(reps=30):
    """back-to-back calls, one event after each (bench_sweeps.py's method: the device-side interval per call)"""
    for _ in range(3):
        ci.convolution_resample(kin, kout, v, kernel="b13")
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        ci.convolution_resample(kin, kout, v, kernel="b13")
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[0], ts[len(ts) // 2]


for env in ({"CIB200_RS_MMA": "0", "CIB200_RS_NO_TABLE_CACHE": "1"}, {"CIB200_RS_NO_TABLE_CACHE": "1"}, {}, {"CIB200_RS_NO_TABLE_CACHE": "1"}, {},
            {"CIB200_NO_PDL": "1"}, {"CIB200_RS_NCONS": "11"}):
    for k in ("CIB200_RS_MMA", "CIB200_RS_NTEAMS", "CIB200_RS_NB", "CIB200_RS_NCONS", "CIB200_NO_PDL", "CIB200_RS_BPI", "CIB200_RS_NO_TABLE_CACHE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b, m = best()
    print(f"{str(env):40s} best {b:.4f} ms  median {m:.4f} ms", flush=True)
