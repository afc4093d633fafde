set -x
python -m pytest tests/test_ppo.py tests/test_learner_gemm.py -m gpu -q 2>&1 | tail -12 | cut -c1-400
python tools/learner_time.py highest 2>&1 | tail -2
