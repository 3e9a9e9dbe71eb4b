import sys, torch
sys.path.insert(0, "/root/repo")
import bench, convitp_b200 as ci
knots, vals = bench.make_grid(512)
for kern in ("b13", "b11", "b9"):
    itp = ci.convolution_interpolation(knots, vals, kernel=kern, bc="poly")
    itp.itp.set_profiling(True)
    for nq in (250_000_000, 1_000_000_000):
        q = [torch.rand(nq, dtype=torch.float64, device="cuda") for _ in range(3)]
        o = torch.empty(nq, dtype=torch.float64, device="cuda")
        best = 1e30
        for _ in range(3):
            itp.eval(*q, out=o); torch.cuda.synchronize()
            best = min(best, sum(itp.itp.last_pass_ms().values()))
        print(kern, nq, itp.itp.last_strategy(), round(best, 2), "ms", f"{nq/best*1e3:.3e} evals/s", flush=True)
        del q, o
    itp.itp.close(); del itp; torch.cuda.empty_cache()
