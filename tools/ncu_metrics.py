#!/usr/bin/env python
"""One JSON record per ncu `--set full` report: what bench.py's roofline quotes (dram traffic, fp32 flops per launch) and
the occupancy / issue / stall figures DESIGN.md cites. Output is merged into profiles/<tag>_metrics.json under a key
`<kernel>|<task>|<terrain>|<num_envs>|dr<0/1>` that bench.py looks up for the configuration it is timing.

    python tools/ncu_metrics.py gpurun_out/r02c_warp.ncu-rep "pgtt_env_kernel<OP_STEP>|stairs|level1|4096|dr0" profiles/r02_metrics.json
"""
import csv, io, json, subprocess, sys
from pathlib import Path


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u, v = rows[0], rows[1], rows[2]
    return {n: (val, unit) for n, unit, val in zip(h, u, v)}


def num(m, name, scale_units=True):
    val, unit = m[name]
    x = float(val)
    if scale_units:
        x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6}.get(unit, 1.0)
    return x


def record(rep):
    m = raw(rep)
    cyc = num(m, "gpc__cycles_elapsed.max")
    per_cycle = {k: num(m, f"smsp__sass_thread_inst_executed_op_{k}_pred_on.sum.per_cycle_elapsed") for k in ("fadd", "fmul", "ffma")}
    flops = (per_cycle["fadd"] + per_cycle["fmul"] + 2.0 * per_cycle["ffma"]) * cyc
    peak_per_cycle = num(m, "derived__sm__sass_thread_inst_executed_op_ffma_pred_on_x2")
    stalls = {k.split("issue_stalled_")[1].split("_per_issue_active")[0].replace("_per_warp_active", ""): float(v[0]) for k, v in m.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")} if any("issue_stalled" in k for k in m) else {}
    return {
        "report": Path(rep).name, "kernel_name": m["Kernel Name"][0], "duration_us": num(m, "gpu__time_duration.sum") * 1e6,
        "grid": int(float(m["launch__grid_size"][0])), "block": int(float(m["launch__block_size"][0])),
        "registers_per_thread": int(float(m["launch__registers_per_thread"][0])),
        "dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
        "fp32_flops": flops, "fp32_flops_note": "(fadd + fmul + 2 ffma) thread instructions, predicated on, per launch",
        "fp32_frac_of_peak_under_ncu": flops / cyc / peak_per_cycle,
        "warp_instructions": num(m, "smsp__inst_executed.sum"),
        "warps_active_pct": num(m, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "stall_cycles_per_issue": {k: round(v, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v >= 0.05},
    }


if __name__ == "__main__":
    rep, key, dst = sys.argv[1], sys.argv[2], Path(sys.argv[3])
    d = json.loads(dst.read_text()) if dst.exists() else {}
    d[key] = record(rep)
    dst.write_text(json.dumps(d, indent=1, sort_keys=True) + "\n")
    print(json.dumps(d[key], indent=1))
