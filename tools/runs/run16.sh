set -x
timeout 600 python -m pytest tests/test_ppo.py tests/test_mlp_native.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
python tools/sgd_step_time.py 2>&1 | tail -1
python tools/learner_time.py highest 2>&1 | tail -1
