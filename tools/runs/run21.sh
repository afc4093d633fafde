set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 20 > gpurun_out/r02f_bench_2gpu.json 2> gpurun_out/r02f_bench_2gpu.err; tail -3 gpurun_out/r02f_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 -m phase_guided_terrain_traversal_b200.train --num_envs 8192 --batch_size 512 --terrain_file level07 --num_timesteps 3000000 --num_evals 4 2>&1 | grep -v Warning | tail -8
