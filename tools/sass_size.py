"""Static code-size report: SASS instructions per kernel, attributed to source functions via -lineinfo.

Development aid for instruction-cache work (the env kernels are I$-bound when their hot loop exceeds
the ~32 KB L1.5 instruction cache): `python tools/sass_size.py [kernel-substring]`.
"""
import bisect, collections, re, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "phase_guided_terrain_traversal_b200" / "csrc"


def func_table():
    tabs = {}
    for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.cu")) + list(CSRC.glob("*.h")):
        starts = []
        for i, t in enumerate(open(p), 1):
            m = re.match(r"^(?:DEV_NOINLINE|DEV|static|template|__global__|__device__)[^;]*?\b([A-Za-z_0-9]+)\s*\(", t)
            if m and not t.startswith(" "):
                starts.append((i, m.group(1)))
        tabs[p.name] = starts
    return tabs


def main():
    lib = sys.argv[2] if len(sys.argv) > 2 else str(CSRC / "libpgtt_b200.so")
    pat = sys.argv[1] if len(sys.argv) > 1 else "ILi0"
    tabs = func_table()
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, capture_output=True)
        cub = list(Path(d).glob("*.cubin"))[0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout
    cur_fn, cur_src = None, ("?", 0)
    per_kernel = collections.Counter()
    per_func = collections.defaultdict(collections.Counter)
    for line in dis.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", line)
        if m:
            cur_fn = m.group(1); cur_src = ("?", 0); continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            cur_src = (Path(m.group(1)).name, int(m.group(2))); continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur_fn:
            per_kernel[cur_fn] += 1
            f, ln = cur_src
            s = tabs.get(f, [])
            k = bisect.bisect_right([a for a, _ in s], ln) - 1
            per_func[cur_fn][(f, s[k][1] if k >= 0 else "?")] += 1
    for k, n in sorted(per_kernel.items(), key=lambda kv: kv[1]):
        print(f"{n:7d} instr {n * 16 / 1024:7.1f} KB  {k}")
    for k in per_kernel:
        if pat in k:
            print(f"\n== {k}")
            for (f, fn), n in per_func[k].most_common(45):
                print(f"  {n:6d} {n * 16 / 1024:6.1f} KB  {f}:{fn}")


if __name__ == "__main__":
    main()
