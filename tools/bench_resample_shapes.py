import sys, os, numpy as np, torch
sys.path.insert(0, "/root/repo")
import convitp_b200 as ci
rng = np.random.default_rng(3)
for kernel, nin, nout in (("b13", 128, 256), ("b5", 128, 256), ("a4", 128, 256), ("b13", 256, 256), ("b7", 200, 300), ("b13", 256, 128)):
    kin = tuple(np.linspace(0, 6.28, nin) for _ in range(3)); kout = tuple(np.linspace(0, 6.28, nout) for _ in range(3))
    v = torch.from_numpy(rng.random((nin,) * 3)).cuda()
    for _ in range(3): ci.convolution_resample(kin, kout, v, kernel=kernel)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        ci.convolution_resample(kin, kout, v, kernel=kernel); ev[i + 1].record()
    torch.cuda.synchronize()
    print(os.environ.get("CIB200_RESAMPLE_LIGHT", "0"), kernel, nin, nout, f"{min(ev[i].elapsed_time(ev[i+1]) for i in range(10)):.4f} ms", flush=True)
