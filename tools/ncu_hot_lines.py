#!/usr/bin/env python
"""Top source lines of one kernel in an ncu report by executed warp-instructions (and their stall samples).
    python tools/ncu_hot_lines.py <report.ncu-rep> <mangled kernel name> [n]"""
import csv, io, sys
from collections import defaultdict
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent))
from ncu_summary import ncu, sass_line_table, CSRC

rep, mangled = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
hi_ = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
h = src[hi_]; rows = [r for r in src[hi_ + 1:] if len(r) == len(h)]
cs, ci = h.index("# Samples"), h.index("Instructions Executed")
lines = sass_line_table(CSRC / "libpgtt_b200.so", mangled)
base = min(int(r[0], 16) for r in rows)
agg = defaultdict(lambda: [0.0, 0.0, 0])
for r in rows:
    f, ln = lines.get(int(r[0], 16) - base, ("?", 0))
    a = agg[(f, ln)]; a[0] += float(r[cs] or 0); a[1] += float(r[ci] or 0); a[2] += 1
tot = sum(a[1] for a in agg.values()); tots = sum(a[0] for a in agg.values())
txt = {}
print(f"total warp-instructions {tot:.0f}, samples {tots:.0f}, sass {len(rows)}")
for (f, ln), a in sorted(agg.items(), key=lambda x: -x[1][1])[:n]:
    if f not in txt:
        try: txt[f] = (CSRC / f).read_text().splitlines()
        except Exception: txt[f] = []
    t = txt[f][ln - 1].strip()[:100] if 0 < ln <= len(txt[f]) else ""
    print(f"{f}:{ln:5d} inst% {100 * a[1] / tot:5.2f} samp% {100 * a[0] / tots:5.2f} sass {a[2]:4d} | {t}")
