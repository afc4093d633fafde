python tools/sgd_step_time.py 2>&1 | tail -2
PGTT_PAR=0 python tools/sgd_step_time.py 2>&1 | tail -1
PGTT_NATIVE_MLP=layers python tools/sgd_step_time.py 2>&1 | tail -1
PGTT_NATIVE_MLP=0 python tools/sgd_step_time.py 2>&1 | tail -1
python tools/learner_time.py highest 2>&1 | tail -1
