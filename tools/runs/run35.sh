set -x
timeout 600 python -m pytest tests/test_mlp_native.py tests/test_ppo.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
python tools/mlp_time.py 2>&1 | grep rows
python tools/sgd_step_time.py 2>&1 | tail -1
