set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 100 --warmup 20 --cpu-seconds 3 > gpurun_out/r02c_bench_quick.json 2> gpurun_out/r02c_bench_quick.err; tail -3 gpurun_out/r02c_bench_quick.err
