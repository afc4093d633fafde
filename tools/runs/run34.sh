set -x
timeout 600 python -m phase_guided_terrain_traversal_b200.train --num_envs 4096 --batch_size 256 --terrain_file level1 --num_timesteps 160000000 --num_evals 17 --out gpurun_out/r02j_policy_level1_1gpu 2>&1 | grep -E "steps|saved" | tail -20 > gpurun_out/r02j_train_1gpu.log
cat gpurun_out/r02j_train_1gpu.log
timeout 200 python -m phase_guided_terrain_traversal_b200.evaluate --policy gpurun_out/r02j_policy_level1_1gpu --terrain_file level1 2>&1 | tail -1 > gpurun_out/r02j_eval_level1_1gpu.log
cat gpurun_out/r02j_eval_level1_1gpu.log
