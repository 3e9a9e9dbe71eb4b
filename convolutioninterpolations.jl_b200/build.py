"""Builds libcib200.so (sm_100a) in-tree with nvcc.  Usage: python build.py [--force] [-v] [--experiments]

--experiments adds -DCIB200_EXPERIMENTS: only such a build reads the CIB200_* tuning variables from the environment
(cib200_internal.cuh tune_env); the default build ignores the environment."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SO = os.path.join(HERE, "libcib200.so")
SOURCES = ["cib200_api.cu", "cib200_eval.cu", "cib200_sweep.cu", "cib200_binned.cu", "cib200_evalnd.cu", "cib200_integral.cu", "cib200_smooth.cu", "cib200_resample.cu", "cib200_integral_nd.cu", "cib200_lanequery.cu", "cib200_cellmma.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# CIB200_EXTRA_NVCC_FLAGS / CIB200_SO_NAME: A/B builds of kernel variants during development (tools/ab_build.sh)
SO = os.path.join(HERE, os.environ.get("CIB200_SO_NAME", "libcib200.so"))
FLAGS = [*os.environ.get("CIB200_EXTRA_NVCC_FLAGS", "").split(),
         "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2", "--fmad=true",
         "-I", os.path.join(ROOT, "include"), "-I", CSRC]


def _deps():
    out = [os.path.join(ROOT, "include", "cib200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build(force=False, verbose=False, experiments=False):
    if not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in _deps()):
        return SO
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        o = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(o)
        cmd = [NVCC, *FLAGS, *(["-DCIB200_EXPERIMENTS"] if experiments else []), "-c", os.path.join(CSRC, s), "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {s}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv or "--experiments" in sys.argv, verbose="-v" in sys.argv,
                experiments="--experiments" in sys.argv))
