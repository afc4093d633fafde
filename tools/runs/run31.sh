for F in 0 1 0 1; do echo "fuse=$F"; PGTT_FUSE_TASK=$F python tools/step_time.py stairs 4096 level1 200 2>&1 | grep -E "back|flushed"; done
PGTT_FUSE_TASK=1 timeout 900 python -m pytest tests/test_kernel_parity.py tests/test_policy_rollout.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
