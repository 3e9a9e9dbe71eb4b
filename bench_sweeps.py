#!/usr/bin/env python3
"""Secondary measurements (not the driver's contract line): BASELINE configs 3 and 4 - the separable sweeps -
device-resident, CUDA events, against the separable-sweep byte model of SURVEY.md 8d.  One JSON line per config."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import convitp_b200 as ci  # noqa: E402

PEAK = 6556.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass



def julia_order(a):
    """CUDA tensor indexed like `a` whose MEMORY is column-major (axis 0 contiguous) - how a Julia array lies in memory.
    The host mirror hands such a tensor to the library as it is; a C-ordered tensor would first be transposed by a
    torch copy kernel inside the timed region (0.1 ms for 64^4)."""
    import numpy as np
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a.T)).cuda()
    return t.permute(*reversed(range(t.dim())))


def timeit(fn, warmup=3, iters=30):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[0], ts[len(ts) // 2]


def graph_ms(fn, calls=10):
    """the same call captured in a CUDA graph and replayed: device time per call without the host side (Python mirror,
    planning, launches); None when the capture fails"""
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for _ in range(calls):
                    fn()
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / calls)
        return best
    except Exception:  # noqa: BLE001
        return None


def main():
    rng = np.random.default_rng(3)
    out = []
    only = sys.argv[1] if len(sys.argv) > 1 else 'all'
    # C3: 3-D resample b13 128^3 -> 256^3 (+ derivative pass)
    kin = tuple(np.linspace(0, 2 * np.pi, 128) for _ in range(3))
    kout = tuple(np.linspace(0, 2 * np.pi, 256) for _ in range(3))
    v = julia_order(rng.random((128, 128, 128)))
    for der in ((0, (1, 0, 0)) if only in ('all', 'resample') else ()):
        best, med = timeit(lambda: ci.convolution_resample(kin, kout, v, kernel="b13", derivative=der))
        gms = None  # (the resample call does not capture: see DESIGN 5.2c for its device time from the ncu launch list)
        mb = (128 ** 3 + 256 * 128 ** 2) + (256 * 128 ** 2 + 256 ** 2 * 128) + (256 ** 2 * 128 + 256 ** 3)
        bytes_ = 8 * mb
        out.append({"config": f"C3 resample b13 128^3->256^3 derivative={der}", "ms_best": best, "ms_median": med, "ms_graph_replay": gms,
                    "value": 256 ** 3 / (best * 1e-3), "unit": "output pts/s",
                    "roofline": {"bound": "hbm", "achieved": bytes_ / (best * 1e-3) / 1e9, "peak": PEAK, "unit": "GB/s",
                                 "frac": bytes_ / (best * 1e-3) / 1e9 / PEAK, "model": "sum over passes (read+write) = 352.3 MB"}})
    # C4: 4-D smooth 40^4 and 64^4, B = 0.05
    for n in ((40, 64) if only in ('all', 'smooth') else ()):
        ks = tuple(np.linspace(0, 2 * np.pi, n) for _ in range(4))
        v4 = julia_order(rng.random((n,) * 4))
        best, med = timeit(lambda: ci.convolution_smooth(ks, v4, 0.05))
        gms = graph_ms(lambda: ci.convolution_smooth(ks, v4, 0.05))
        bytes_ = 4 * 2 * 8 * n ** 4
        out.append({"config": f"C4 smooth {n}^4 B=0.05", "ms_best": best, "ms_median": med, "ms_graph_replay": gms,
                    "value": n ** 4 / (best * 1e-3), "unit": "grid pts/s",
                    "roofline": {"bound": "hbm", "achieved": bytes_ / (best * 1e-3) / 1e9, "peak": PEAK, "unit": "GB/s",
                                 "frac": bytes_ / (best * 1e-3) / 1e9 / PEAK, "model": "4 passes x 16 B per grid point",
                                 "note": "40^4 (20 MB) is L2-resident: frac may exceed 1"}})
    if only != 'all':
        for o in out:
            print(json.dumps(o), flush=True)
        return
    # C1 / C2 point evaluation (latency / cache bound: reported, not graded against HBM)
    x = np.linspace(0, 1, 1000)
    it1 = ci.convolution_interpolation(x, np.sin(2 * np.pi * x))
    q1 = torch.rand(1_000_000, dtype=torch.float64, device="cuda")
    best, med = timeit(lambda: it1.eval(q1))
    out.append({"config": "C1 1-D b5 n=1000, 1e6 queries", "ms_best": best, "value": 1e6 / (best * 1e-3), "unit": "evals/s"})
    x2 = np.linspace(0, 2 * np.pi, 200)
    it2 = ci.convolution_interpolation((x2, x2), np.sin(x2)[:, None] * np.cos(x2)[None, :], derivative=(1, 0), extrap="line")
    q2 = [(torch.rand(10_000_000, dtype=torch.float64, device="cuda") * 1.2 - 0.1) * 2 * np.pi for _ in range(2)]
    best, med = timeit(lambda: it2.eval(*q2))
    out.append({"config": "C2 2-D b5 200^2 derivative=(1,0) Line, 1e7 queries", "ms_best": best,
                "value": 1e7 / (best * 1e-3), "unit": "evals/s"})
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
