set -x
PGTT_KERNEL=warp ncu --set full --clock-control none --import-source on -k regex:pgtt_env_kernel -s 40 -c 1 -f -o gpurun_out/r02c_warp python tools/kernel_times.py stairs 4096 level1 10 > gpurun_out/prof3.log 2>&1
