set -x
for W in 28 14; do PGTT_WARPS_PER_BLOCK=$W PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 100; done
PGTT_KERNEL=warp PGTT_SYNC_MASK=0 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_SYNC_MASK=2047 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_SYNC_MASK=256 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_SYNC_MASK=384 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=warp PGTT_SYNC_MASK=386 python tools/kernel_times.py stairs 4096 level1 100
for N in 2048 8192; do PGTT_KERNEL=warp python tools/kernel_times.py stairs $N level1 100; done
