set -x
timeout 600 python -m pytest tests/test_ppo.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
python tools/train_step_phases.py 2>&1 | grep world
PGTT_N=8192 PGTT_LEVEL=level07 python tools/train_step_phases.py 2>&1 | grep world
