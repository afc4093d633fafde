# ncu evidence for profiles/ (round 2): launch list of the bench command + one full capture per kernel of the step
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 240 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02e_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_env_kernel -s 6 -c 1 -f -o gpurun_out/r02e_warp python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_task_kernel -s 6 -c 1 -f -o gpurun_out/r02e_task python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_quad_kernel -s 6 -c 1 -f -o gpurun_out/r02e_quad python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --num-envs 8192 --terrain level07 --dr 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_gemm_kernel -s 10 -c 3 -f -o gpurun_out/r02e_gemm python tools/learner_gemm_time.py > gpurun_out/prof6_gemm.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 600 --csv --log-file gpurun_out/r02e_learner_launches.csv python tools/learner_time.py highest > gpurun_out/prof6_learner.log 2>&1
for M in 0 2 258; do PGTT_KERNEL=warp PGTT_SYNC_MASK=$M python tools/kernel_times.py stairs 4096 level1 100 2>&1 | grep back; done
ls -la gpurun_out/r02e_*
