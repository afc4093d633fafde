set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp python tools/stage_trace.py 4096 level1 2>&1 | tail -14
