set -x
timeout 900 python -m pytest tests/test_ppo.py tests/test_policy_rollout.py tests/test_mlp_native.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
python tools/learner_time.py highest 2>&1 | tail -1
python tools/train_step_phases.py 2>&1 | tail -1
