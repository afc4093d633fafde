set -x
PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 100
for M in 0 2 256 258 386 2047; do PGTT_KERNEL=warp PGTT_SYNC_MASK=$M python tools/kernel_times.py stairs 4096 level1 60 | head -1; done
PGTT_KERNEL=warp ncu --set full --clock-control none --import-source on -k regex:pgtt_env_kernel -s 40 -c 1 -f -o gpurun_out/r02b_warp python tools/kernel_times.py stairs 4096 level1 10 > gpurun_out/prof2.log 2>&1
PGTT_KERNEL=warp ncu --set full --clock-control none --import-source on -k regex:pgtt_task_kernel -s 40 -c 1 -f -o gpurun_out/r02b_task python tools/kernel_times.py stairs 4096 level1 10 >> gpurun_out/prof2.log 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
