#!/usr/bin/env python
"""Summarise an ncu report for profiles/: headline metrics (raw page) and per-function shares of the
warp-stall samples / executed instructions (source page; needs -lineinfo and --import-source on).

    python tools/ncu_summary.py gpurun_out/r01b_step.ncu-rep [launch-index]

Function attribution: SASS addresses of the source page -> file:line through `nvdisasm --print-line-info`
of the in-tree library (must be the build that was profiled) -> enclosing function of csrc/*.
"""
from __future__ import annotations

import csv
import io
import re
import subprocess
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CSRC = ROOT / "phase_guided_terrain_traversal_b200" / "csrc"

RAW_KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def function_table():
    """(file name, line) -> enclosing function, from the kernel sources."""
    out = {}
    pat = re.compile(r"^(?:DEV_NOINLINE|DEV|static|__global__|template|struct).*?\b([A-Za-z_][A-Za-z0-9_]*)\s*[({]")
    for f in sorted(CSRC.glob("*.cu*")) + sorted(CSRC.glob("*.h")):
        cur = "?"
        for n, line in enumerate(f.read_text().splitlines(), 1):
            mm = pat.match(line)
            if mm and not line.rstrip().endswith(";"):
                cur = mm.group(1)
            out[(f.name, n)] = cur
    return out


def sass_line_table(lib: Path, kernel_mangled: str):
    """instruction offset -> (file, line) of one kernel, from `nvdisasm --print-line-info` of the built library."""
    import tempfile
    tmp = Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(lib)], cwd=tmp, capture_output=True)
    table = {}
    for cubin in tmp.glob("*.cubin"):
        dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", str(cubin)], capture_output=True, text=True).stdout
        inside, cur = False, ("?", 0)
        for line in dis.splitlines():
            if line.startswith("//---------------------"):
                inside = (".text." + kernel_mangled + " ") in line + " "
                continue
            if not inside:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (Path(m.group(1)).name, int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
            if m:
                table[int(m.group(1), 16)] = cur
    return table


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    mangled = sys.argv[3] if len(sys.argv) > 3 else "_Z15pgtt_env_kernelILi0EEv10LaunchArgs"
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, data = raw[0], raw[1], raw[2:]
    print(f"## raw metrics (launch {idx} of {len(data)}: {data[idx][hdr.index('Kernel Name')]})")
    for k in RAW_KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:70s} {data[idx][i]:>16s} {units[i]}")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    st = sorted(((float(data[idx][i]), hdr[i][len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for i in stall_cols), reverse=True)
    print("stall cycles per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in st if v >= 0.05))

    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    tables, cur = [], None
    for r in src:
        if r and r[0] == "Address":
            cur = {"hdr": r, "rows": []}
            tables.append(cur)
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    t = tables[min(idx, len(tables) - 1)]
    h = t["hdr"]
    c_samp, c_inst = h.index("# Samples"), h.index("Instructions Executed")
    stall_names = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
    stall_idx = [h.index(s) for s in stall_names]
    lines = sass_line_table(CSRC / "libpgtt_b200.so", mangled)
    fn_of = function_table()
    base = min(int(r[0], 16) for r in t["rows"])
    agg = defaultdict(lambda: defaultdict(float))
    tot = defaultdict(float)

    def num(x):
        try:
            return float(x)
        except ValueError:
            return 0.0
    for r in t["rows"]:
        loc = lines.get(int(r[0], 16) - base, ("?", 0))
        fn = (loc[0], fn_of.get(loc, "?"))
        agg[fn]["samples"] += num(r[c_samp]); agg[fn]["inst"] += num(r[c_inst]); agg[fn]["sass"] += 1
        tot["samples"] += num(r[c_samp]); tot["inst"] += num(r[c_inst]); tot["sass"] += 1
        for s, i in zip(stall_names, stall_idx):
            agg[fn][s] += num(r[i])
    print(f"\n## per-function shares (stall samples {tot['samples']:.0f}, warp-instructions executed {tot['inst']:.0f}, SASS instructions {tot['sass']:.0f})")
    print(f"{'file':20s} {'function':22s} {'samples%':>8s} {'inst%':>7s} {'sass':>6s}  top stalls (share of the function's samples)")
    for fn, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:36]:
        tops = sorted(((a[s], s[6:]) for s in stall_names), reverse=True)[:3]
        ssum = max(sum(a[s] for s in stall_names), 1.0)
        print(f"{fn[0]:20s} {fn[1]:22s} {100 * a['samples'] / tot['samples']:8.1f} {100 * a['inst'] / tot['inst']:7.1f} {a['sass']:6.0f}  "
              + ", ".join(f"{n} {100 * v / ssum:.0f}%" for v, n in tops))


if __name__ == "__main__":
    main()
