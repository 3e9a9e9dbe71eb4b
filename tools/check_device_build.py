"""Host-built vs device-built coefficient grids (cib200_build_coefs vs cib200_build_coefs_device): max difference and
where it sits.  python tools/check_device_build.py [n ...]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convitp_b200 as ci

for n in [int(a) for a in sys.argv[1:]] or [64, 128, 256]:
    rng = np.random.default_rng(n)
    g = np.linspace(0, 1, n)
    v = np.sin(2 * np.pi * g)[:, None, None] * np.cos(2 * np.pi * g)[None, :, None] * np.sin(4 * np.pi * g + 1)[None, None, :]
    v = v + 0.01 * rng.uniform(-1, 1, v.shape)
    for kernel, bc in (("b5", "poly"), ("b5", "detect"), ("b13", "poly")):
        a = ci.convolution_interpolation((g, g, g), v, kernel=kernel, bc=bc).itp.coefs
        b = ci.convolution_interpolation((g, g, g), v, kernel=kernel, bc=bc, coefs_on="device").itp.coefs
        d = np.abs(a - b)
        i = np.unravel_index(d.argmax(), d.shape)
        print(n, kernel, bc, "max diff", d.max(), "at", i, "of", a.shape, "values", a[i], b[i], flush=True)
