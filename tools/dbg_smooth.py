import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CASES = {"1d64": (64,), "2d64": (64, 64), "2d40": (40, 40), "3d40": (40, 40, 40), "4d40": (40,) * 4, "4d64": (64,) * 4, "3d_50_64_30": (50, 64, 30)}
if len(sys.argv) > 1 and sys.argv[1] in CASES:
    import convitp_b200 as ci
    from oracle import oracle as O
    shape = CASES[sys.argv[1]]
    rng = np.random.default_rng(1)
    ks = tuple(np.linspace(0, 2 * np.pi, n) for n in shape)
    v = rng.random(shape) + 3.0
    B = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
    got = ci.convolution_smooth(ks, v, B)
    ref = O.convolution_smooth(ks, v, B)
    print(sys.argv[1], os.environ.get("CIB200_SMOOTH_NEW"), os.environ.get("CIB200_NO_PDL"), "max err", float(np.abs(got - ref).max()), flush=True)
else:
    for env in ({"CIB200_SMOOTH_NEW": "1"}, {"CIB200_SMOOTH_NEW": "2"}, {}, {"CIB200_NO_PDL": "1"}):
        for c in CASES:
            e = dict(os.environ, **env)
            r = subprocess.run([sys.executable, __file__, c], env=e, capture_output=True, text=True, timeout=300)
            out = (r.stdout + r.stderr).strip().splitlines()
            print(env, c, "rc", r.returncode, " | ".join(out[-3:])[:400], flush=True)
