#!/usr/bin/env python3
"""Timing of convolution_smooth (B = 0.05) on N^4 Float64 arrays, device-resident, column-major input: best / median
of 20 event-timed calls, and the host time per call (no synchronisation) to tell a host-bound loop from a GPU-bound one.
Tuning variables are read only by an --experiments build."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci  # noqa: E402

for n in (40, 64):
    ks = tuple(np.linspace(0, 2 * np.pi, n) for _ in range(4))
    v4 = torch.from_numpy(np.random.default_rng(3).random((n,) * 4)).cuda().permute(3, 2, 1, 0)
    for _ in range(3):
        ci.convolution_smooth(ks, v4, 0.05)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
    t0 = time.perf_counter()
    ev[0].record()
    for i in range(20):
        ci.convolution_smooth(ks, v4, 0.05)
        ev[i + 1].record()
    host = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(20))
    print(f"n={n} best {ts[0]*1e3:.1f} us median {ts[10]*1e3:.1f} us host/call {host*1e6:.1f} us  env "
          + " ".join(f"{k[7:]}={v}" for k, v in os.environ.items() if k.startswith("CIB200_")), flush=True)
    # the same call captured in a CUDA graph and replayed: device time without the host side of the call
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            ci.convolution_smooth(ks, v4, 0.05)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(10):
                    ci.convolution_smooth(ks, v4, 0.05)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"   graph replay: {e0.elapsed_time(e1) * 100:.1f} us per call", flush=True)
    except Exception as ex:  # noqa: BLE001
        print("   graph capture failed:", str(ex)[:200], flush=True)
