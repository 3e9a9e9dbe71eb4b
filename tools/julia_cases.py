#!/usr/bin/env python3
"""Seeded parity cases for a run of the REAL reference (tools/julia_dump.jl), and the checker for its outputs.

    python tools/julia_cases.py write DIR [--full]   inputs of BASELINE configs C1-C5 (+ the two convolution_smooth
                                                      boundary quirks of SURVEY.md 3.5) as raw little-endian Float64
    julia --project=<reference> tools/julia_dump.jl DIR
    python tools/julia_cases.py check DIR [--gpu]     out.f64 / coefs.f64 of every case against the oracle (and,
                                                      with --gpu, against the CUDA library through the C-ABI)

Knot vectors are written explicitly (Julia's range() is not numpy.linspace).  --full writes
config 5 on its 512^3 grid (1 GB of values); the default is the same field on 128^3.
Case directory: spec.txt (key=value), knots_<d>.f64, values.f64 (column-major), x_<d>.f64 | knots_out_<d>.f64."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 0xC1B200


def _write_case(root, name, spec, knots, values, xs=None, knots_out=None):
    d = os.path.join(root, name)
    os.makedirs(d, exist_ok=True)
    nd = values.ndim
    spec = dict(spec, ndim=nd, n=",".join(str(s) for s in values.shape))
    with open(os.path.join(d, "spec.txt"), "w") as f:
        for k, v in spec.items():
            f.write(f"{k}={v}\n")
    for i, k in enumerate(knots):
        np.asarray(k, dtype="<f8").tofile(os.path.join(d, f"knots_{i}.f64"))
    np.asarray(values, dtype="<f8").ravel(order="F").tofile(os.path.join(d, "values.f64"))
    for i, x in enumerate(xs or []):
        np.asarray(x, dtype="<f8").tofile(os.path.join(d, f"x_{i}.f64"))
    for i, k in enumerate(knots_out or []):
        np.asarray(k, dtype="<f8").tofile(os.path.join(d, f"knots_out_{i}.f64"))


def _csv(v, n):
    v = (v,) * n if isinstance(v, (str, int)) else tuple(v)
    return ",".join(str(x) for x in v)


def c5_field(n, rng):
    g = np.linspace(0.0, 1.0, n)
    a = np.sin(2 * np.pi * g) + 0.5 * np.cos(6 * np.pi * g + 0.3)
    b = np.cos(2 * np.pi * g) + 0.3 * np.sin(10 * np.pi * g)
    c = np.sin(4 * np.pi * g + 1.0)
    v = np.empty((n, n, n), order="F")
    for k in range(n):
        v[:, :, k] = a[:, None] * b[None, :] * c[k] + 0.01 * rng.uniform(-1, 1, (n, n))
    return g, v


def write_cases(root, full=False, small=False):
    """small=True: tiny shapes of every kind (the stub pipeline test); otherwise BASELINE's shapes."""
    os.makedirs(root, exist_ok=True)
    interp = dict(kind="interp", extrap="throw", subgrid="cubic")
    # C1: 1-D :b5, n = 1000, 10^6 queries, :detect and :poly
    rng = np.random.default_rng(SEED + 1)
    n1, q1 = (60, 2000) if small else (1000, 1_000_000)
    x = np.linspace(0, 1, n1)
    v = np.sin(2 * np.pi * x) + 0.1 * rng.uniform(-1, 1, n1)
    q = rng.uniform(0, 1, q1)
    for bc in ("detect", "poly"):
        _write_case(root, f"c1_{bc}", dict(interp, kernel="b5", derivative="0", bc=bc), (x,), v, (q,))
    # C2: 2-D :b5 200^2, derivative (1, 0), Line extrapolation, ~31 % of the queries outside
    rng = np.random.default_rng(SEED + 2)
    n2, q2 = (30, 3000) if small else (200, 1_000_000)
    x = np.linspace(0, 2 * np.pi, n2)
    v = np.sin(x)[:, None] * np.cos(x)[None, :] + 0.05 * rng.uniform(-1, 1, (n2, n2))
    q = rng.uniform(-0.1 * 2 * np.pi, 1.1 * 2 * np.pi, (2, q2))
    _write_case(root, "c2_line", dict(interp, kernel="b5,b5", derivative="1,0", bc="detect", extrap="line", subgrid="cubic,cubic"),
                (x, x), v, tuple(q))
    # C3: resample :b13 128^3 -> 256^3, derivative 0 and (1, 0, 0)
    rng = np.random.default_rng(SEED + 3)
    n3, m3 = (20, 33) if small else (128, 256)
    ki = tuple(np.linspace(0, 2 * np.pi, n3) for _ in range(3))
    ko = tuple(np.linspace(0, 2 * np.pi, m3) for _ in range(3))
    g = np.meshgrid(*ki, indexing="ij")
    v = np.sin(g[0]) * np.cos(g[1]) * np.sin(g[2]) + 0.01 * rng.uniform(-1, 1, (n3,) * 3)
    for name, dv in (("c3_d000", "0,0,0"), ("c3_d100", "1,0,0")):
        _write_case(root, name, dict(kind="resample", kernel="b13,b13,b13", derivative=dv, bc="detect", subgrid="cubic,cubic,cubic"),
                    ki, v, knots_out=ko)
    # C4: smooth B = 0.05, 40^4 (+5 offset variant exposes the mean-centring quirk) and the 1-D quirk probes
    rng = np.random.default_rng(SEED + 4)
    n4 = 26 if small else 40
    nd4 = 2 if small else 4
    ks = tuple(np.linspace(0, 2 * np.pi, n4) for _ in range(nd4))
    g = np.meshgrid(*ks, indexing="ij")
    v = np.ones((n4,) * nd4)
    for i, gi in enumerate(g):
        v = v * (np.sin(gi) if i % 2 == 0 else np.cos(gi))
    v = v + 0.1 * rng.standard_normal(v.shape)
    _write_case(root, "c4_smooth", dict(kind="smooth", B="0.05"), ks, v)
    _write_case(root, "c4_smooth_plus5", dict(kind="smooth", B="0.05"), ks, v + 5.0)
    xq = np.linspace(0, 2 * np.pi, 200)
    _write_case(root, "smooth_quirk_cos", dict(kind="smooth", B="0.05"), (xq,), np.cos(xq))
    _write_case(root, "smooth_quirk_cos_plus5", dict(kind="smooth", B="0.05"), (xq,), np.cos(xq) + 5.0)
    # C5: 3-D :b5 on the bench's field, :poly and :detect, 2^20 uniformly random queries
    rng = np.random.default_rng(SEED + 5)
    n5 = 24 if small else (512 if full else 128)
    gk, v = c5_field(n5, rng)
    q = rng.uniform(0, 1, (3, 4000 if small else 1 << 20))
    for bc in ("poly", "detect"):
        _write_case(root, f"c5_{bc}", dict(interp, kernel="b5,b5,b5", derivative="0,0,0", bc=bc, subgrid="cubic,cubic,cubic"),
                    (gk, gk, gk), v, tuple(q))
    # b13 and derivatives in 3-D with Flat extrapolation
    q = rng.uniform(-0.05, 1.05, (3, 4000 if small else 200_000))
    nb = 24 if small else 64
    gk, v = c5_field(nb, rng)
    _write_case(root, "b13_d100_flat", dict(interp, kernel="b13,b13,b13", derivative="1,0,0", bc="detect", extrap="flat",
                                            subgrid="cubic,cubic,cubic"), (gk, gk, gk), v, tuple(q))


def read_case(d):
    spec = dict(ln.strip().split("=", 1) for ln in open(os.path.join(d, "spec.txt")) if ln.strip())
    nd = int(spec["ndim"])
    n = tuple(int(x) for x in spec["n"].split(","))
    rd = lambda f: np.fromfile(os.path.join(d, f), dtype="<f8")
    case = {"spec": spec, "ndim": nd, "knots": tuple(rd(f"knots_{i}.f64") for i in range(nd)),
            "values": rd("values.f64").reshape(n, order="F")}
    if spec["kind"] == "interp":
        case["xs"] = tuple(rd(f"x_{i}.f64") for i in range(nd))
    if spec["kind"] == "resample":
        case["knots_out"] = tuple(rd(f"knots_out_{i}.f64") for i in range(nd))
    if os.path.exists(os.path.join(d, "out.f64")):
        case["out"] = rd("out.f64")
    if os.path.exists(os.path.join(d, "coefs.f64")):
        case["coefs"] = rd("coefs.f64")
    return case


def run_case(impl, case):
    """impl: the oracle module or the CUDA host mirror (same function names and keywords).  Returns (flat output, scale,
    ghost-padded grid or None)."""
    sp, nd = case["spec"], case["ndim"]
    one = lambda s, conv=str: tuple(conv(x) for x in s.split(","))
    if sp["kind"] == "interp":
        itp = impl.convolution_interpolation(case["knots"] if nd > 1 else case["knots"][0], case["values"],
                                             kernel=one(sp["kernel"]), derivative=one(sp["derivative"], int), bc=sp["bc"],
                                             extrap=sp["extrap"], subgrid=one(sp["subgrid"]))
        out = np.asarray(itp.eval(*case["xs"]))
        inner = getattr(itp, "itp", itp)
        coefs = np.asarray(inner.coefs)
        scale = np.abs(coefs).max() * np.prod([(1.0 / (k[1] - k[0])) ** d for k, d in zip(case["knots"], one(sp["derivative"], int))])
        if sp["extrap"] == "line":  # f_b + (f_b - f_in)/eps * dist amplifies any rounding difference by 2 dist/eps
            lo = np.array([k[0] for k in case["knots"]])[:, None]
            hi = np.array([k[-1] for k in case["knots"]])[:, None]
            X = np.stack(case["xs"])
            dist = np.sqrt(((X - np.clip(X, lo, hi)) ** 2).sum(axis=0))
            scale = scale * (1.0 + 2.0 * dist / (min(k[1] - k[0] for k in case["knots"]) / 100.0))
        return out.ravel(), scale, coefs.ravel(order="F")
    if sp["kind"] == "resample":
        out = impl.convolution_resample(case["knots"], case["knots_out"], case["values"], kernel=one(sp["kernel"]),
                                        derivative=one(sp["derivative"], int), bc=sp["bc"], subgrid=one(sp["subgrid"]))
        scale = np.abs(case["values"]).max() * np.prod([(1.0 / (k[1] - k[0])) ** d for k, d in zip(case["knots"], one(sp["derivative"], int))])
        return np.asarray(out).ravel(order="F"), scale, None
    out = impl.convolution_smooth(case["knots"], case["values"], float(sp["B"]))
    return np.asarray(out).ravel(order="F"), np.abs(case["values"]).max(), None


def check_cases(root, impls, tol=1e-12):
    """Compares out.f64 (and coefs.f64) of every case that has them with each implementation.  Returns
    {case: {impl: max relative error}}; raises AssertionError on the first violation."""
    report = {}
    for name in sorted(os.listdir(root)):
        d = os.path.join(root, name)
        if not os.path.exists(os.path.join(d, "out.f64")):
            continue
        case = read_case(d)
        report[name] = {}
        for label, impl in impls.items():
            got, scale, coefs = run_case(impl, case)
            err = float((np.abs(got - case["out"]) / np.maximum(np.abs(case["out"]), scale)).max())
            report[name][label] = err
            assert err <= tol, f"{name}: {label} differs from the reference output by {err:.3e} (relative)"
            if coefs is not None and "coefs" in case:
                cerr = float(np.abs(coefs - case["coefs"]).max() / np.abs(case["coefs"]).max())
                report[name][label + "_coefs"] = cerr
                assert cerr <= tol, f"{name}: {label} ghost-padded grid differs from the reference's by {cerr:.3e}"
    return report


def main():
    if len(sys.argv) < 3 or sys.argv[1] not in ("write", "check"):
        raise SystemExit(__doc__)
    root = sys.argv[2]
    if sys.argv[1] == "write":
        write_cases(root, full="--full" in sys.argv, small="--small" in sys.argv)
        print(f"wrote {len(os.listdir(root))} cases under {root}")
        return
    from oracle import oracle as O
    impls = {"oracle": O}
    if "--gpu" in sys.argv:
        import convitp_b200 as ci
        impls["cuda"] = ci
    rep = check_cases(root, impls)
    if not rep:
        raise SystemExit(f"no case under {root} has an out.f64: run tools/julia_dump.jl first")
    for name, r in rep.items():
        print(name, " ".join(f"{k}={v:.2e}" for k, v in r.items()))
    src = os.path.join(root, "JULIA_VERSION")
    print("parity PINNED against", open(src).read().strip() if os.path.exists(src) else "an unnamed producer")


if __name__ == "__main__":
    main()
