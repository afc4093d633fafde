set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 30 > gpurun_out/r02j_bench_8gpu.json 2> gpurun_out/r02j_bench_8gpu.err; tail -2 gpurun_out/r02j_bench_8gpu.err
timeout 200 $TR --master-port 29522 bench.py --gpus 8 --steps 200 --warmup 30 --terrain curriculum --dr 1 --no-extras --no-cpu-baseline > gpurun_out/r02j_bench_8gpu_curriculum.json 2> gpurun_out/r02j_bench_8gpu_curriculum.err
timeout 400 $TR --master-port 29523 -m phase_guided_terrain_traversal_b200.train --num_envs 65536 --batch_size 2048 --terrain_file level07 --num_timesteps 120000000 --num_evals 7 --out gpurun_out/r02j_policy_level07 2>&1 | grep -v Warning | tail -18 > gpurun_out/r02j_train_8gpu.log
cat gpurun_out/r02j_train_8gpu.log
timeout 200 python -m phase_guided_terrain_traversal_b200.evaluate --policy gpurun_out/r02j_policy_level07 --terrain_file level07 2>&1 | tail -1 > gpurun_out/r02j_eval_level07.log
cat gpurun_out/r02j_eval_level07.log
