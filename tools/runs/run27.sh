set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300
python bench.py > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; tail -3 gpurun_out/r02g_bench.err
