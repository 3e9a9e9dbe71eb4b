"""Single-process multi-GPU host-buffer throughput (cib200_itp_eval_multi): 512^3 b5 grid, pinned host queries.
python tools/bench_multi.py [nq=4e8]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci

nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 400_000_000
n = 512
rng = np.random.default_rng(1)
g = np.linspace(0, 1, n)
v = rng.random((n, n, n))
hx = [torch.rand(nq, dtype=torch.float64).pin_memory() for _ in range(3)]
hout = torch.empty(nq, dtype=torch.float64).pin_memory()
xs = [h.numpy() for h in hx]
ndev = torch.cuda.device_count()
for devs in [list(range(k)) for k in (1, 2, 4, 8) if k <= ndev]:
    itp = ci.convolution_interpolation((g, g, g), v, kernel="b5", bc="poly", devices=devs)
    itp.eval(*xs, out=hout.numpy())
    t0 = time.perf_counter()
    itp.eval(*xs, out=hout.numpy())
    dt = time.perf_counter() - t0
    print(f"devices={devs}: {nq / dt:.3e} evals/s ({dt * 1e3:.1f} ms for {nq} host-resident queries)", flush=True)
    itp.itp.close()
