#!/usr/bin/env python3
"""Driver for ncu captures of the sweep kernels: one C3 resample (b13 128^3 -> 256^3) and one C4 smooth (64^4, 40^4),
device-resident, after one warm-up call each.  Usage: ncu ... python tools/prof_sweeps.py [resample|smooth|all]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

def julia_order(a):
    """CUDA tensor indexed like `a` whose MEMORY is column-major (axis 0 contiguous) - how a Julia array lies in memory.
    The host mirror hands such a tensor to the library as it is; a C-ordered tensor would first be transposed by a
    torch copy kernel inside the timed region (0.1 ms for 64^4)."""
    import numpy as np
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    return t.permute(*reversed(range(t.dim())))


what = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(3)
if what in ("resample", "all"):
    kin = tuple(np.linspace(0, 2 * np.pi, 128) for _ in range(3))
    kout = tuple(np.linspace(0, 2 * np.pi, 256) for _ in range(3))
    v = julia_order(rng.random((128, 128, 128)))
    for _ in range(2):
        ci.convolution_resample(kin, kout, v, kernel="b13")
    torch.cuda.synchronize()
if what in ("smooth", "all"):
    for n in (64, 40):
        ks = tuple(np.linspace(0, 2 * np.pi, n) for _ in range(4))
        v4 = julia_order(rng.random((n,) * 4))
        for _ in range(2):
            ci.convolution_smooth(ks, v4, 0.05)
        torch.cuda.synchronize()
