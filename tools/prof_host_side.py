import cProfile, pstats, sys, os, io
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import convitp_b200 as ci
from bench_sweeps import julia_order
rng = np.random.default_rng(3)
n = 40
ks = tuple(np.linspace(0, 2 * np.pi, n) for _ in range(4))
v4 = julia_order(rng.random((n,) * 4))
for _ in range(5):
    ci.convolution_smooth(ks, v4, 0.05)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(200):
    ci.convolution_smooth(ks, v4, 0.05)
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host per call us", (t1 - t0) / 200 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    ci.convolution_smooth(ks, v4, 0.05)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
print(s.getvalue()[:3500])
kin = tuple(np.linspace(0, 2 * np.pi, 128) for _ in range(3))
kout = tuple(np.linspace(0, 2 * np.pi, 256) for _ in range(3))
v = julia_order(rng.random((128, 128, 128)))
for _ in range(5):
    ci.convolution_resample(kin, kout, v, kernel="b13")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    ci.convolution_resample(kin, kout, v, kernel="b13")
t1 = time.perf_counter()
torch.cuda.synchronize()
print("resample host per call us", (t1 - t0) / 200 * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    ci.convolution_resample(kin, kout, v, kernel="b13")
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(14)
print(s.getvalue()[:3500])
