#!/usr/bin/env python3
"""Summarise an .ncu-rep (one `ncu --set full` capture) into profiles/<name>.json + .txt.

    python tools/ncu_summary.py gpurun_out/prof_eval_r01.ncu-rep profiles/r01_eval_gather --evals 20000000
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_fp64.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--evals", type=float, default=None, help="units of work processed by the captured launch")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = {"source": a.rep, "note": a.note, "kernels": []}
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")], "metrics": {}}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                k["metrics"][key] = {"value": r[i], "unit": units[i]}
        m = k["metrics"]

        def val(key):
            v = m.get(key)
            if not v:
                return None
            return float(v["value"].replace(",", "")) * UNIT.get(v["unit"], 1)
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        if rd is not None and wr is not None:
            k["dram_bytes_per_launch"] = rd + wr
            if a.evals:
                k["dram_bytes_per_eval"] = (rd + wr) / a.evals
        out["kernels"].append(k)
    json.dump(out, open(a.out + ".json", "w"), indent=1)
    with open(a.out + ".txt", "w") as f:
        f.write(f"# {a.rep}\n# {a.note}\n")
        for k in out["kernels"]:
            f.write(f"\n## {k['kernel']}\n")
            for key, v in k["metrics"].items():
                f.write(f"{key:75s} {v['value']:>22s} {v['unit']}\n")
            if "dram_bytes_per_eval" in k:
                f.write(f"{'dram bytes per evaluation':75s} {k['dram_bytes_per_eval']:22.1f} byte\n")
    print(open(a.out + ".txt").read())


if __name__ == "__main__":
    main()
