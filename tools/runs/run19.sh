set -x
timeout 600 python -m pytest tests/test_ppo.py tests/test_mlp_native.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-300
python tools/mlp_time.py 2>&1 | grep rows
PGTT_MLP_WIDE=1 python tools/mlp_time.py 2>&1 | grep rows
python tools/sgd_step_time.py 2>&1 | tail -1
PGTT_MLP_WIDE=1 python tools/sgd_step_time.py 2>&1 | tail -1
python tools/train_step_phases.py 2>&1 | tail -1
