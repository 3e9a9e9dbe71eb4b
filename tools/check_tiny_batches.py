import sys, numpy as np
sys.path.insert(0, "/root/repo")
import convitp_b200 as ci
rng = np.random.default_rng(1)
worst = 0.0
for kernel in ("b5", "b7", "b13"):
    for shape in ((20, 20, 20), (36, 22, 28), (18, 40, 26)):
        k = tuple(np.linspace(0, 1, n) for n in shape)
        v = rng.random(shape)
        gb = ci.convolution_interpolation(k, v, kernel=kernel, eval_mode="binned", extrap="flat")
        gg = ci.convolution_interpolation(k, v, kernel=kernel, eval_mode="gather", extrap="flat")
        ncell = np.prod([n - 1 for n in shape])
        for nq in (1, 7, 33, 1000, int(0.45 * ncell), int(1.3 * ncell), int(3.5 * ncell), int(9 * ncell)):
            q = rng.uniform(-0.01, 1.01, (3, nq))
            a, b = gb.eval(*q), gg.eval(*q)
            err = np.abs(a - b).max() / np.abs(v).max()
            worst = max(worst, err)
            assert err < 1e-13, (kernel, shape, nq, err, gb.itp.last_strategy())
        print(kernel, shape, "ok", gb.itp.last_strategy(), flush=True)
print("worst", worst)
