#!/usr/bin/env python3
"""Aggregate the pc-sampling data of an .ncu-rep (captured with --import-source on) into buckets of consecutive SASS
instructions: share of stall samples, executed warp instructions and the dominant opcodes / stall reasons per bucket.

    python tools/ncu_hotspots.py gpurun_out/x.ncu-rep [bucket=40] [min_pct=1.5]
"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
MINP = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
for n, s in enumerate(secs):
    e = secs[n + 1] if n + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]
    data = [r for r in rows[s + 2:e] if len(r) == len(hdr)]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[si] or 0) for r in data) or 1
    texec = sum(int(r[ii] or 0) for r in data)
    print(f"== [{n}] {rows[s][1][:90]}  samples {tot}  warp-instr {texec}  static {len(data)}")
    for b in range(0, len(data), B):
        blk = data[b:b + B]
        smp = sum(int(r[si] or 0) for r in blk)
        if smp < tot * MINP / 100:
            continue
        ie = sum(int(r[ii] or 0) for r in blk)
        ops = collections.Counter()
        for r in blk:
            t = r[1].split()
            op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
            ops[op.split(".")[0]] += 1
        st = collections.Counter()
        for i in stall_cols:
            st[hdr[i].replace("stall_", "")] += sum(int(r[i] or 0) for r in blk)
        print(f"  [{b:5d}..{b + B:5d}) samples {100 * smp / tot:5.1f}%  exec {100 * ie / max(texec, 1):5.1f}%  "
              f"ops {dict(ops.most_common(3))}  stalls {dict(st.most_common(3))}")
