set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for M in 258 256 384 386; do PGTT_KERNEL=warp PGTT_SYNC_MASK=$M python tools/kernel_times.py stairs 4096 level1 100 2>&1 | grep back; done
PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level07 100 1
