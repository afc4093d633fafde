#!/usr/bin/env python
"""Per-source-line instruction / stall-sample counts of one kernel from an ncu report (see ncu_summary.py).
    python tools/ncu_lines.py <report> <mangled kernel> <file name> <first line> <last line>"""
import csv, io, sys
from collections import defaultdict
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent))
from ncu_summary import ncu, sass_line_table, CSRC

rep, mangled, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
hi_ = [i for i, r in enumerate(src) if r and r[0] == "Address"][0]
h = src[hi_]; rows = [r for r in src[hi_ + 1:] if len(r) == len(h)]
cs, ci = h.index("# Samples"), h.index("Instructions Executed")
lines = sass_line_table(CSRC / "libpgtt_b200.so", mangled)
base = min(int(r[0], 16) for r in rows)
agg = defaultdict(lambda: [0.0, 0.0, 0])
for r in rows:
    f, ln = lines.get(int(r[0], 16) - base, ("?", 0))
    if f == fname and lo <= ln <= hi:
        a = agg[ln]; a[0] += float(r[cs] or 0); a[1] += float(r[ci] or 0); a[2] += 1
text = (CSRC / fname).read_text().splitlines()
tot = sum(float(r[ci] or 0) for r in rows)
for ln in sorted(agg):
    a = agg[ln]
    print(f"{ln:5d} samples {a[0]:6.0f} inst% {100 * a[1] / tot:5.2f} sass {a[2]:4d} | {text[ln - 1].strip()[:110]}")
