set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for W in 14 7; do PGTT_WARPS_PER_BLOCK=$W PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 100; done
for N in 2048 4144 8192; do PGTT_KERNEL=warp python tools/kernel_times.py stairs $N level1 100; done
PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level07 100 1
PGTT_KERNEL=warp PGTT_SYNC_MASK=0 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_SYNC_MASK=2047 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_QUAD_FULLSCAN=1 python tools/kernel_times.py stairs 4096 level1 100
