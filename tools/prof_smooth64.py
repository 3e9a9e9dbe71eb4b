#!/usr/bin/env python3
"""ncu driver: one convolution_smooth of a 64^4 (or N^4) Float64 array, B = 0.05, device-resident, after a warm-up."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ks = tuple(np.linspace(0, 2 * np.pi, n) for _ in range(4))
a = np.random.default_rng(3).random((n,) * 4)
t = torch.from_numpy(a).cuda()
v4 = t.permute(3, 2, 1, 0)  # memory = column-major of the indexed array
for _ in range(2):
    ci.convolution_smooth(ks, v4, 0.05)
torch.cuda.synchronize()
