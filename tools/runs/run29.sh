TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/train_step_phases.py 2>&1 | grep "world"
PGTT_N=8192 PGTT_BS=512 PGTT_LEVEL=level07 timeout 300 $TR --master-port 29542 tools/train_step_phases.py 2>&1 | grep "world"
PGTT_N=8192 PGTT_BS=256 PGTT_LEVEL=level07 timeout 300 python tools/train_step_phases.py 2>&1 | grep "world"
