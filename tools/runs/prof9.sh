# final-build captures (round 2): quad physics kernel at config[2], the learner GEMM, launch list of the SGD steps
set -x
ncu --set full --clock-control none --import-source on -k regex:pgtt_quad_kernel -s 6 -c 1 -f -o gpurun_out/r02k_quad python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --num-envs 8192 --terrain level07 --dr 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:pgtt_bgemm_kernel -s 44 -c 11 -f -o gpurun_out/r02k_bgemm python tools/mlp_time.py > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 500 --csv --log-file gpurun_out/r02k_learner_launches.csv python tools/learner_time.py highest > gpurun_out/prof9_learner.log 2>&1
python tools/sgd_step_time.py 2>&1 | tail -1
python tools/mlp_time.py 2>&1 | grep rows
ls -la gpurun_out/r02k_*
