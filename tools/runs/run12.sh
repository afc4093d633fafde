set -x
python -m pytest tests -m gpu -q 2>&1 | tail -15 | cut -c1-400
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; tail -3 gpurun_out/r02e_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02e_bench_reference.json 2>/dev/null
python tools/learner_time.py highest 2>&1 | tail -2
