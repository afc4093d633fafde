set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/e_tests.log
python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
for qw in 1 2 4 8; do
  PGTT_KERNEL=quad PGTT_QUAD_WARPS=$qw python bench.py --no-cpu-baseline --steps 100 > gpurun_out/e_quad4096_qw$qw.json 2>/dev/null
done
for qw in 2 4; do
  PGTT_KERNEL=quad PGTT_QUAD_WARPS=$qw python bench.py --no-cpu-baseline --steps 100 --num-envs 2048 > gpurun_out/e_quad2048_qw$qw.json 2>/dev/null
done
python bench.py --no-cpu-baseline --steps 100 --num-envs 2048 > gpurun_out/e_warp2048.json 2>/dev/null
cat gpurun_out/e_tests.log
