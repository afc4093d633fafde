"""`python -m phase_guided_terrain_traversal_b200.train` - the call sequence of training/train.py:102-268 over the
B200-native env, rollout collector and PPO learner (same flags as training/train.py:285-296).

    python -m phase_guided_terrain_traversal_b200.train --task_name stairs --terrain_file level1 --num_envs 4096 \
        --num_timesteps 10000000 [--out policy_out]
    torchrun --nproc-per-node 8 -m phase_guided_terrain_traversal_b200.train --num_envs 65536 --batch_size 2048 ...

Under torchrun every rank owns `num_envs / world` envs (index sharding, sharding.py); gradients and observation
statistics are all-reduced over NCCL.
"""
from __future__ import annotations

import argparse
import functools
import os
import time

import numpy as np


def run_training(args):
    import torch
    import torch.distributed as dist
    from . import ppo, registry, sharding, terrain, wrapper
    from .go2 import joystick, joystick_pgtt, randomize, randomize_simple
    from .go2.configs import baseline_config, default_config

    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    env_name = "Go2"
    joy, cfg_fn = (joystick_pgtt, default_config) if args.method == "pgtt" else (joystick, baseline_config)                     # train.py:111-114,119-122
    registry.register_environment(env_name, functools.partial(joy.Joystick, task=args.task_name, device=local), cfg_fn)          # train.py:115-116
    env_cfg = cfg_fn()                                                                                       # train.py:119-129
    env_cfg.command_config.u_max = [0.6, 0.6, 1.0]
    env_cfg.command_config.u_min = [-0.6, -0.6, -1.0]
    env_cfg.gait_freq = [1, 3]
    env = registry.load(env_name, config=env_cfg)
    if args.task_name == "stairs":                                                                                   # train.py:165-170
        matrix = terrain.load_terrain(args.terrain_file)
        registry._randomizer[env_name] = functools.partial(randomize.domain_randomize, terrain_matrix=matrix)
    else:
        registry._randomizer[env_name] = randomize_simple.domain_randomize
    cfg = ppo.PPOConfig(num_timesteps=args.num_timesteps, episode_length=env_cfg.episode_length, num_minibatches=args.num_minibatches,
                        discounting=args.discount, learning_rate=args.learning_rate, num_envs=args.num_envs, batch_size=args.batch_size,
                        seed=args.seed, matmul_precision=args.matmul_precision,
                        global_advantage_norm=bool(args.global_advantage_norm))                                                                                # train.py:135-161
    keys = sharding.shard_keys(args.seed, args.num_envs, rank, world)
    t0 = time.time()

    x_data, y_data, y_dataerr, linvels, angvels = [], [], [], [], []                                                # train.py:182-186
    scales = env_cfg.reward_config.scales

    def progress(num_steps, metrics):                                                                                # train.py:198-229
        """Called on rank 0 with the evaluator's metrics; returns the reference's convergence verdict (True = stop)."""
        x_data.append(num_steps)
        y_data.append(metrics["eval/episode_reward"])
        y_dataerr.append(metrics["eval/episode_reward_std"])
        vel_tracking_per = metrics["eval/episode_reward/tracking_lin_vel"] / (scales.tracking_lin_vel * env_cfg.episode_length)
        ang_tracking_per = metrics["eval/episode_reward/tracking_ang_vel"] / (scales.tracking_ang_vel * env_cfg.episode_length)
        linvels.append(vel_tracking_per)
        angvels.append(ang_tracking_per)
        print(f"steps {num_steps:>12d}  eval reward {y_data[-1]:.3f} +- {y_dataerr[-1]:.3f}  Lin vel {vel_tracking_per:.3f}  ang vel {ang_tracking_per:.3f}  "
              f"episode length {metrics['eval/avg_episode_length']:.0f}  training reward/step {metrics.get('training/reward_per_step', float('nan')):.5f}  "
              f"{num_steps / max(time.time() - t0, 1e-9):,.0f} env-steps/s", flush=True)
        if len(y_data) >= 2 and y_data[-1] != 0:                                                                      # termination criteria, train.py:224-228
            rel = abs((y_data[-1] - y_data[-2]) / y_data[-1])
            if vel_tracking_per > env_cfg.vel_percentage and ang_tracking_per > env_cfg.vel_percentage and rel <= 0.005:
                return True
            if rel <= 0.001:
                return True
        return False

    restore = None
    if args.checkpoint_folder is not None:                                                                          # train.py:246-256
        from pathlib import Path
        ck = Path(args.checkpoint_folder)
        if ck.is_dir():      # like get_max_numbered_folder (train.py:271-283): the entry with the largest numeric name
            nums = [p for p in ck.iterdir() if p.name.isdigit()]
            if not nums:
                raise SystemExit(f"--checkpoint_folder {ck}: no numbered checkpoints inside")
            ck = max(nums, key=lambda p: int(p.name))
        restore = ck
        if rank == 0:
            print(f"Restoring from checkpoint: {ck}")
    ckdir = None
    if args.save_checkpoints:
        from pathlib import Path
        ckdir = Path(args.save_checkpoints)
        if rank == 0:
            ckdir.mkdir(parents=True, exist_ok=True)

    def policy_params_fn(num_steps, trainer):                                                                        # train.py:189-196
        if rank == 0 and ckdir is not None:
            trainer.save(ckdir / f"{num_steps}")

    trainer = ppo.train(environment=env, eval_env=registry.load(env_name, config=env_cfg) if rank == 0 else None,                    # train.py:242-263
                        wrap_env_fn=wrapper.wrap_for_brax_training, randomization_fn=registry.get_domain_randomizer(env_name),
                        rng_keys=keys, cfg=cfg, progress_fn=progress, policy_params_fn=policy_params_fn, restore_checkpoint_path=restore,
                        num_evals=args.num_evals, num_eval_envs=args.num_eval_envs)
    if rank == 0 and args.plots:                                                                                      # train.py:267-270
        from pathlib import Path
        pl = Path(args.plots)
        pl.mkdir(parents=True, exist_ok=True)
        for name, arr in (("mean", y_data), ("std", y_dataerr), ("lin_vel", linvels), ("anf_vel", angvels), ("steps", x_data)):
            np.save(pl / f"{name}{args.index}", np.asarray(arr))
    if rank == 0 and args.out:
        trainer.save(args.out)                                                                                       # model.save_params, train.py:266
        print(f"saved {args.out} (layout of deploy/policy_net.py:6-33)")
    if world > 1:
        # a captured graph that contains NCCL kernels must be gone before its communicator is torn down
        trainer._graph = None
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return trainer


def main():
    p = argparse.ArgumentParser(description="Train PPO on the B200-native GO2 PGTT env")
    p.add_argument("--method", type=str, default="pgtt")
    p.add_argument("--task_name", type=str, default="stairs")
    p.add_argument("--terrain_file", type=str, default="level1")
    p.add_argument("--num_envs", type=int, default=4096)
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--discount", type=float, default=0.97)
    p.add_argument("--learning_rate", type=float, default=3e-4)
    p.add_argument("--num_minibatches", type=int, default=32)
    p.add_argument("--num_timesteps", type=int, default=1)
    p.add_argument("--num_evals", type=int, default=31, help="evaluations over the run, brax schedule (the first one before training)")
    p.add_argument("--num_eval_envs", type=int, default=128, help="environments of the in-training evaluator (brax default)")
    p.add_argument("--index", type=int, default=32, help="suffix of the arrays written to --plots")
    p.add_argument("--plots", type=str, default=None, help="folder for the evaluation curves (mean / std / lin_vel / anf_vel, train.py:267-270)")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--out", type=str, default=None)
    p.add_argument("--checkpoint_folder", type=str, default=None, help="resume from a policy pickle, or from the highest-numbered one in a folder")
    p.add_argument("--save_checkpoints", type=str, default=None, help="folder that receives one pickle per training step, named by env-step count")
    p.add_argument("--global_advantage_norm", type=int, default=1,
                   help="1: advantages normalised with moments all-reduced over all ranks (NCCL; the north-star variant), 0: per-rank minibatch (brax)")
    p.add_argument("--matmul_precision", type=str, default="highest", help="learner GEMMs: highest (fp32, reference) or high (TF32)")
    args = p.parse_args()
    if args.method not in ("pgtt", "baseline"):
        raise SystemExit("--method must be pgtt or baseline")
    run_training(args)


if __name__ == "__main__":
    main()
