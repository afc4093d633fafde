set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 4 --steps 100 --warmup 20 > gpurun_out/r02g_bench_4gpu.json 2> gpurun_out/r02g_bench_4gpu.err; tail -2 gpurun_out/r02g_bench_4gpu.err
