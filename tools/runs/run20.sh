set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300
python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -3 gpurun_out/r02f_bench.err
python tools/train_step_phases.py 2>&1 | tail -1
