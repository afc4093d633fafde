set -x
python -m pytest tests/test_ppo.py -m gpu -q -k "evaluation or trainer_runs" 2>&1 | tail -12 | cut -c1-300
for W in 14 7; do PGTT_KERNEL=warp PGTT_WARPS_PER_BLOCK=$W python tools/kernel_times.py stairs 4096 level1 100 2>&1 | grep back; done
