set -x
PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 200
python -m pytest tests/test_kernel_parity.py -m gpu -q -k "reset_and_step or stage_by_stage or baseline" 2>&1 | tail -3
