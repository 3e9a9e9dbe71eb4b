"""BASELINE config 2 (2-D b5 200^2, 10^7 queries, derivative=(1,0), Line extrapolation) for ncu / timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci

x2 = np.linspace(0, 2 * np.pi, 200)
extrap = sys.argv[1] if len(sys.argv) > 1 else "line"
it2 = ci.convolution_interpolation((x2, x2), np.sin(x2)[:, None] * np.cos(x2)[None, :], derivative=(1, 0), extrap=extrap)
span = (1.2, -0.1) if extrap != "throw" else (1.0, 0.0)
q2 = [(torch.rand(10_000_000, dtype=torch.float64, device="cuda") * span[0] + span[1]) * 2 * np.pi for _ in range(2)]
for _ in range(3):
    it2.eval(*q2)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    it2.eval(*q2)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"C2 extrap={extrap}: {ms:.3f} ms, {1e7 / ms / 1e-3:.3e} evals/s")
