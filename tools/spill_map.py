#!/usr/bin/env python
"""Where does a kernel touch local memory? Counts LDL/STL SASS instructions per source function (nvdisasm line info).

    python tools/spill_map.py [_Z16pgtt_quad_kernelILi0EEv10LaunchArgs]
"""
import re, subprocess, sys, tempfile
from collections import Counter
from pathlib import Path
sys.path.insert(0, str(Path(__file__).parent))
from ncu_summary import function_table, CSRC

kernel = sys.argv[1] if len(sys.argv) > 1 else "_Z16pgtt_quad_kernelILi0EEv10LaunchArgs"
tmp = Path(tempfile.mkdtemp())
subprocess.run(["cuobjdump", "-xelf", "all", str(CSRC / "libpgtt_b200.so")], cwd=tmp, capture_output=True)
ft = function_table()
cnt, lines = Counter(), Counter()
for cubin in tmp.glob("*.cubin"):
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", str(cubin)], capture_output=True, text=True).stdout
    inside, cur = False, ("?", 0)
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            inside = kernel in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (Path(m.group(1)).name, int(m.group(2)))
            continue
        m = re.search(r"\b(LDL|STL)(\.\w+)*\b", ln)
        if m:
            fn = ft.get(cur, "?")
            cnt[(fn, m.group(1))] += 1
            lines[(cur[0], cur[1], m.group(1))] += 1
print("per function:")
for (fn, op), n in cnt.most_common(40):
    print(f"  {fn:28s} {op} {n}")
print("per line (top 40):")
for (f, l, op), n in lines.most_common(40):
    print(f"  {f}:{l} {op} {n}")
