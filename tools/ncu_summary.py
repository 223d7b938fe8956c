#!/usr/bin/env python
"""Summarise one `ncu --set full` capture (a .ncu-rep) of a kernel as JSON for profiles/.

  python tools/ncu_summary.py REPORT.ncu-rep OUT.json --points N [--sass-sha FILE] [--note TEXT]

Executed FP64 flops = 2*DFMA + DMUL + DADD (smsp__sass_thread_inst_executed_op_d*_pred_on),
DRAM traffic = dram__bytes_read.sum + dram__bytes_write.sum, both per launch.  `sass_sha256`
identifies the library build the capture was taken from (bench.py refuses to use a capture
whose hash differs from the library it times).  Debug/measurement tooling only.
"""
import argparse
import csv
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report"); ap.add_argument("out")
    ap.add_argument("--points", type=int, required=True)
    ap.add_argument("--sass-sha", default=None)
    ap.add_argument("--note", default="")
    ap.add_argument("--workload", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    col = {n: i for i, n in enumerate(h)}

    def get(name, scale=True):
        i = col[name]
        x = float(v[i].replace(",", ""))
        return x * UNIT.get(u[i], 1.0) if scale else x

    def thread_inst(op):
        # the .sum is not always collected; per_cycle_elapsed.sum * elapsed cycles is
        name = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
        if name in col:
            return get(name)
        return get(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed") * get("smsp__cycles_elapsed.avg")

    t_ms = get("gpu__time_duration.sum")
    dfma, dmul, dadd = thread_inst("dfma"), thread_inst("dmul"), thread_inst("dadd")
    flop = 2 * dfma + dmul + dadd
    stalls = {}
    for n, i in col.items():
        if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio") and "not_issued" not in n:
            stalls[n[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = float(v[i])
    tot = sum(stalls.values()) or 1.0
    out = {
        "source": f"ncu --set full --clock-control none, one launch of {v[col['Kernel Name']]} ({a.report})",
        "workload": a.workload, "note": a.note,
        "sass_sha256": open(a.sass_sha).read().split()[0] if a.sass_sha else None,
        "points_per_launch": a.points,
        "grid_size": get("launch__grid_size", False), "registers_per_thread": get("launch__registers_per_thread", False),
        "gpu_time_ms": t_ms,
        "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
        "dram_bytes_per_point": (get("dram__bytes_read.sum") + get("dram__bytes_write.sum")) / a.points,
        "thread_inst_dfma": dfma, "thread_inst_dmul": dmul, "thread_inst_dadd": dadd,
        "executed_fp64_flop": flop, "executed_fp64_flop_per_point": flop / a.points,
        "executed_fp64_flop_per_s_under_ncu": flop / (t_ms * 1e-3),
        "fp64_pipe_pct_of_peak_active": get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "warps_active_pct": get("sm__warps_active.avg.pct_of_peak_sustained_active", False),
        "warp_inst_executed": get("smsp__inst_executed.sum", False),
        "local_load_inst": get("sass__inst_executed_local_loads", False),
        "local_store_inst": get("sass__inst_executed_local_stores", False),
        "register_spill_inst": get("sass__inst_executed_register_spilling", False),
        "l2_hit_rate_pct": get("lts__t_sector_hit_rate.pct", False),
        "warp_stall_pct": {k: round(100 * x / tot, 1) for k, x in sorted(stalls.items(), key=lambda kv: -kv[1])[:10]},
    }
    json.dump(out, open(a.out, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
