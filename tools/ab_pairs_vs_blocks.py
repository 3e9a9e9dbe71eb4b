import os, sys, torch
sys.path.insert(0, "/root/repo")
import bench
import convitp_b200 as ci
knots, vals = bench.make_grid(512)
itp = ci.convolution_interpolation(knots, vals, kernel="b5", bc="poly", eval_mode="binned")
del vals
NQ = 600_000_000
xs = [torch.rand(NQ, dtype=torch.float64, device="cuda") for _ in range(3)]
out = torch.empty(NQ, dtype=torch.float64, device="cuda")
itp.itp.set_profiling(True)
ncell = 511 ** 3
for dens in (1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.4):
    m = int(dens * ncell)
    row = []
    for lq in ("3", "4"):
        os.environ["CIB200_LANEQUERY"] = lq
        os.environ.pop("CIB200_MMA_VARIANT", None)
        best = 1e30
        for _ in range(3):
            itp.eval(*[x[:m] for x in xs], out=out[:m])
            torch.cuda.synchronize()
            best = min(best, sum(itp.itp.last_pass_ms().values()))
        row.append(best)
    print(f"{dens:.1f} per cell ({m:.3e} queries): pairs {row[0]:8.2f} ms   blocks {row[1]:8.2f} ms", flush=True)
