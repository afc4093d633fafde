# ncu evidence for profiles/: launch list of the bench command + one full capture per step-kernel generation
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 160 --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_env_kernel -s 6 -c 1 -f -o gpurun_out/h_warp python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_quad_kernel -s 6 -c 1 -f -o gpurun_out/h_quad python bench.py --steps 4 --warmup 3 --no-cpu-baseline --num-envs 8192 --terrain level07 --dr 1 > /dev/null 2>&1
ls -la gpurun_out/h_*
