set -x
timeout 600 python -m pytest tests/test_ppo.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/train_step_phases.py 2>&1 | grep "world"
timeout 300 $TR --master-port 29552 -m phase_guided_terrain_traversal_b200.train --num_envs 8192 --batch_size 512 --terrain_file level07 --num_timesteps 3000000 --num_evals 3 2>&1 | grep steps | tail -3
