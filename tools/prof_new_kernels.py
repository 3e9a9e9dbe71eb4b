"""Workload for ncu captures of the round-1 late kernels: the device-side constructor passes (pad_lines_axis0_kernel / ghost_lines_kernel, 512^3
b5 grid) and the N-D evaluator (eval_nd_kernel: 4-D b5 on 40^4 and 4-D a4 on 64^4, 2*10^6 random queries each).
    ncu --set full --clock-control none --import-source on -k regex:'pad_lines|ghost_lines|eval_nd' -o gpurun_out/prof_late python tools/prof_new_kernels.py
Without ncu it prints CUDA-event timings of the same launches."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci

L = ci.lib()
dev = torch.device("cuda", 0)
st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


n = 512
vals = torch.rand(n ** 3, dtype=torch.float64, device=dev)
coefs = torch.empty((n + 8) ** 3, dtype=torch.float64, device=dev)
nn = (C.c_int64 * 3)(n, n, n)
hh = (C.c_double * 3)(*[1.0 / (n - 1)] * 3)
kid = (C.c_int32 * 3)(0, 0, 0)
for bc_id, name in ((1, "poly"), (0, "detect")):
    bc = (C.c_int32 * 3)(bc_id, bc_id, bc_id)
    ms = timed(lambda: ci.check(L.cib200_build_coefs_device(3, 0, nn, C.c_void_p(vals.data_ptr()), hh, kid, bc, bc,
                                                            C.c_void_p(coefs.data_ptr()), 1, st)))
    # pass 0: read the values, write lines + ghosts; passes 1, 2 (in place): read the lines, write the ghosts only
    p_ = n + 8
    moved = 8 * (n ** 3 + p_ * n * n + (p_ * n * n + p_ * 8 * n) + (p_ * p_ * n + p_ * p_ * 8))
    print(f"build_coefs_device 512^3 b5 bc={name}: {ms:.3f} ms, {moved / ms / 1e6:.0f} GB/s (bytes read + written by the 3 passes)")
del vals, coefs

rng = np.random.default_rng(5)
for kernel, m in (("b5", 40), ("a4", 64)):
    k = tuple(np.linspace(0, 1, m) for _ in range(4))
    itp = ci.convolution_interpolation(k, rng.random((m,) * 4), kernel=kernel)
    q = [torch.rand(2_000_000, dtype=torch.float64, device=dev) for _ in range(4)]
    ms = timed(lambda: itp.eval(*q))
    taps = {"b5": 10, "a4": 6}[kernel] ** 4
    print(f"eval_nd 4-D {kernel} {m}^4, 2e6 queries: {ms:.3f} ms, {2e6 / (ms * 1e-3):.3e} evals/s, "
          f"{2e6 * taps * 8 / ms / 1e6:.0f} GB/s of stencil reads (gather model)")
