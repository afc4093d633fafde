#!/usr/bin/env python
"""bench.py - env-steps/sec of the batched GO2 PGTT step (BASELINE.json metric) on N B200s.

Workload (N=1): BASELINE config[1] = `Joystick("stairs")` on terrains/level1.npy, 4096 envs, no
dynamics DR (terrain assignment only), synthetic random joystick commands (env-internal
`sample_command`) and U(-1,1) actions, episode wrapper + auto-reset on. One "step" = one wrapped
`env.step` over all envs = ONE launch up to 4144 envs per GPU (warp-per-env physics, 4 x mjx.step, with the task layer - contacts,
ray grid, obs, rewards, wrappers - fused behind it), two above (quad-per-env physics kernel + task kernel). N>1: every rank owns its own 4096 envs (index sharding, no collective
in the data path) -> weak scaling. BASELINE.md section 3: 50 warm-up + 500 timed steps (the defaults).

Timed regions (device time, CUDA events on the launching stream, max over ranks):
  value            K steps, each bracketed by its own event pair (plus one between the two kernels), L2 flushed
                   (256 MiB write) between steps - the state of 4096 envs (~20 MB) would otherwise stay L2-resident;
  value_l2_resident the same K steps back to back with no flush (how a rollout actually runs);
  e2e              K steps through the public API with HOST (pinned) actions: H2D copy of the
                   actions, step, D2H of reward+done, host sync every step;
  rollout          the native PPO collector (policy MLP on tcgen05 -> step, transition slot written by the task kernel);
  config2          BASELINE config[2]: 8192 envs, level07, randomize.py on (its own handle);
  strong           32768 envs in total split by index over the ranks (config[3]'s size): the strong-scaling series;
  train_step       one full PPO training step at the reference hyper-parameters (training/train.py:135-161), fp32.
`roofline` reads the dram traffic and the fp32 flop count of the dominant kernel from the committed ncu record
profiles/*_metrics.json (tools/ncu_metrics.py) for the configuration being timed.
`--impl reference` times the CPU restatement of the reference step (oracle/, fp32, OpenMP over envs,
all host threads) - the reference's own MJX stack is not installable here or on the GPU box
(no jax / mujoco wheels, no network), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import ctypes
import functools
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import numpy as np

METRIC = "env-steps/sec (batched GO2 PGTT step)"
UNIT = "env-steps/s"
# algorithmic bytes per env-step of the step (SURVEY.md 8d / DESIGN.md): 852 B read + 2700 B written
B_ALG = 3552
N_SM, FP32_LANES = 148, 128      # fp32 peak = SMs x lanes x 2 (FMA) x SM clock


def load_kernel_metrics():
    """ncu records committed under profiles/ (tools/ncu_metrics.py), newest round last: key -> record."""
    out = {}
    for f in sorted((ROOT / "profiles").glob("*_metrics.json")):
        try:
            for k, v in json.loads(f.read_text()).items():
                out[k] = dict(v, metrics_file=f"profiles/{f.name}")
        except (OSError, ValueError):
            pass
    return out


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=500)
    p.add_argument("--warmup", type=int, default=50)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--num-envs", type=int, default=4096, help="envs per GPU")
    p.add_argument("--total-envs", type=int, default=0, help="strong-scaling mode: this many envs in total, split by index over the ranks (sharding.shard); overrides --num-envs")
    p.add_argument("--task", default="stairs")
    p.add_argument("--terrain", default="level1", help="level name, or 'curriculum': rank r steps on level{(r mod 10) + 1:02d} (BASELINE config[3])")
    p.add_argument("--dr", type=int, default=0, help="1 = full go2/randomize.py dynamics DR (config 3)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU work budget of the cpu_baseline leg")
    p.add_argument("--no-extras", action="store_true", help="skip the config2 / strong / train_step sub-records")
    p.add_argument("--strong-total", type=int, default=32768, help="total env count of the strong-scaling sub-record")
    return p.parse_args()


def load_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# clocks sampler (NVML), runs during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz, self.ok = [], set(), None, False
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
            self._thr = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU legs (oracle): the ONLY place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
def make_oracle(args, n, cfg, table, seed_offset=0):
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs are meant to use every host thread
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle.oracle import Oracle
    from phase_guided_terrain_traversal_b200 import model as gm, prng
    m = gm.compile_model(args.task, sim_dt=cfg.sim_dt, Kp=cfg.Kp, Kd=cfg.Kd)
    orc = Oracle(m, cfg, n, "f32native")   # gcc -O3 -march=native, built on this host
    try:                                   # libgomp may have read OMP_NUM_THREADS=1 before the line above changed it
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count() or 1)
    except OSError:
        pass
    keys = prng.env_keys(0, n, seed_offset)
    orc.randomize(keys, table if args.task == "stairs" else None, bool(args.dr))
    orc.reset(keys + np.uint32(1))
    return orc


def time_oracle(orc, n, steps, warmup, seed=99):
    g = np.random.default_rng(seed)
    acts = [g.uniform(-1, 1, (n, 12)) for _ in range(8)]
    for i in range(warmup):
        orc.step(acts[i % 8], wrapped=True)
    t0 = time.perf_counter()
    for i in range(steps):
        orc.step(acts[i % 8], wrapped=True)
    return time.perf_counter() - t0


def cpu_baseline(args, cfg, table, budget_s):
    cores = os.cpu_count() or 1
    n = 16 * cores
    orc = make_oracle(args, n, cfg, table)
    t1 = time_oracle(orc, n, 2, 3)
    per_step = t1 / 2
    steps = int(max(5, min(2000, budget_s / max(per_step, 1e-6))))
    t = time_oracle(orc, n, steps, 0)
    return {"value": n * steps / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n} envs x {steps} wrapped steps of the same workload ({args.task}/{args.terrain}, dr={args.dr}), fp32 C oracle (gcc -O3 -march=native), OpenMP over envs, {t:.1f} s"}


def run_reference(args, cfg, table, rank):
    """--impl reference: the CPU restatement on all host threads; each step = a bounded sample of envs."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    probe_n = 16 * cores
    orc = make_oracle(args, probe_n, cfg, table)
    per_env_step = time_oracle(orc, probe_n, 2, 3) / (2 * probe_n)
    total_steps = args.steps + args.warmup
    n = int(max(cores, min(args.num_envs, 60.0 / max(per_env_step * total_steps, 1e-9))))
    n = max(cores, (n // cores) * cores)
    del orc
    orc = make_oracle(args, n, cfg, table)
    t = time_oracle(orc, n, args.steps, max(args.warmup, 3))
    v = n * args.steps / t
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "impl": "reference",
        # the product arm's configuration (same flags -> same dict); what this arm actually stepped is `cpu_baseline.sample`
        "config": workload_config(args, args.num_envs, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"each step = {n} envs on this host's {cores} threads (bounded sample of the {args.num_envs} envs/GPU x {args.gpus} GPU workload; the CPU rate does not "
                                   f"depend on the env count beyond one env per thread), fp32 C restatement of the MJX step, OpenMP over envs"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference MJX/JAX stack not installable (no jax/mujoco wheels, no network): this is the restated CPU oracle, not MJX",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_per_gpu, n_gpus):
    terr = "level01..level10 by rank (curriculum)" if getattr(args, "curriculum", False) else args.terrain
    return {"workload": f"GO2 joystick_pgtt {args.task} terrains/{terr}.npy, {n_per_gpu} envs/GPU x {n_gpus} GPU, "
                        f"{'randomize.py on' if args.dr else 'no DR (terrain assignment only)'}, wrapped step (episode + auto-reset), 4 substeps",
            "num_envs_per_gpu": n_per_gpu, "task": args.task, "terrain": args.terrain, "dr": bool(args.dr),
            "actions": "U(-1,1), pool of 16 pre-generated device buffers", "l2": "flushed between timed steps (256 MiB write); value_l2_resident = back-to-back",
            "launches_per_step": "1 up to 4144 envs per GPU (warp-per-env physics with the task layer fused behind it), 2 above (quad-per-env physics kernel, task kernel)"}


# ------------------------------------------------------------------------------------------------
# product arm helpers
# ------------------------------------------------------------------------------------------------
class Timer:
    """Device timing of wrapped steps of one handle: back to back, or one event triple per step with the L2 flushed in
    between (physics kernel | task kernel split through the development hook that launches one half of pgtt_step)."""

    def __init__(self, torch, dev, abi, pool, barrier, max_over_ranks):
        self.torch, self.dev, self.abi, self.pool, self.barrier, self.max_over_ranks = torch, dev, abi, pool, barrier, max_over_ranks
        self.stream = torch.cuda.current_stream(dev)
        self.part = abi.lib.pgtt_internal_step_part
        self.part.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        self.flush = None

    def resident(self, K):
        torch = self.torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record(self.stream)
        for i in range(K):
            self.abi.step_ptr(self.pool[i % len(self.pool)].data_ptr(), wrapped=True)
        e1.record(self.stream)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def flushed(self, K):
        """-> (total ms of the K steps, mean ms of the physics kernel, mean ms of the task kernel). A handle whose step is ONE launch
        (generation 1, task layer fused behind the physics) is timed through the product call `pgtt_step`: the second figure is then the
        whole fused kernel and the third is 0."""
        torch = self.torch
        if self.flush is None:
            self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        st = ctypes.c_void_p(self.stream.cuda_stream)
        fused = self.abi.step_launches() == 1
        self.barrier()
        for i in range(K):
            self.flush.fill_(i & 0xFF)
            p = self.pool[i % len(self.pool)].data_ptr()
            evs[i][0].record(self.stream)
            if fused:
                self.abi.step_ptr(p, wrapped=True)
                evs[i][1].record(self.stream)
                evs[i][2].record(self.stream)
                continue
            rc = self.part(self.abi.h, p, 1, 0, st)
            evs[i][1].record(self.stream)
            rc |= self.part(self.abi.h, p, 1, 1, st)
            evs[i][2].record(self.stream)
            if rc:
                raise RuntimeError("pgtt_internal_step_part failed")
        self.barrier()
        if fused:
            tot = [e[0].elapsed_time(e[1]) for e in evs]
            return self.max_over_ranks(float(sum(tot))), float(np.mean(tot)), 0.0
        tot = [e[0].elapsed_time(e[2]) for e in evs]
        return self.max_over_ranks(float(sum(tot))), float(np.mean([e[0].elapsed_time(e[1]) for e in evs])), float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))


def make_env(args, cfg, local_rank, N, offset, task, table, dr, seed=0):
    from phase_guided_terrain_traversal_b200 import prng
    from phase_guided_terrain_traversal_b200.go2 import randomize, randomize_simple
    from phase_guided_terrain_traversal_b200.go2.joystick_pgtt import Joystick
    from phase_guided_terrain_traversal_b200.wrapper import wrap_for_brax_training
    keys = prng.env_keys(seed, N, offset=offset)
    env = Joystick(task=task, config=cfg, device=local_rank)
    if task == "stairs":
        rfn = functools.partial(randomize.domain_randomize, rng=keys, terrain_matrix=table, dynamics=bool(dr))
    else:
        rfn = functools.partial(randomize_simple.domain_randomize, rng=keys, dynamics=bool(dr))
    wenv = wrap_for_brax_training(env, episode_length=cfg.episode_length, action_repeat=1, randomization_fn=rfn)
    state = wenv.reset(keys + np.uint32(1))
    return env, wenv, state


def action_pool(torch, dev, N, seed, count=16):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    return [torch.rand((N, 12), generator=g, device=dev) * 2 - 1 for _ in range(count)]


def sub_record(torch, dev, args, cfg, local_rank, rank, world, N, offset, terrain_name, dr, K, W, barrier, max_over_ranks, total_envs):
    """A second handle with its own env count / terrain / DR setting: warm-up, K flushed + K resident steps."""
    from phase_guided_terrain_traversal_b200 import terrain
    table = terrain.load_terrain(terrain_name)
    env, wenv, state = make_env(args, cfg, local_rank, N, offset, "stairs", table, dr, seed=3)
    pool = action_pool(torch, dev, N, 4321 + rank, 8)
    for i in range(W):
        wenv.step(state, pool[i % 8])
    t = Timer(torch, dev, env._abi, pool, barrier, max_over_ranks)
    ms_res = t.resident(K)
    ms_cold, phys_ms, task_ms = t.flushed(K)
    rec = {"value": total_envs * K / (ms_cold * 1e-3), "unit": UNIT, "ms_per_step": ms_cold / K, "value_l2_resident": total_envs * K / (ms_res * 1e-3),
           "ms_per_step_l2_resident": ms_res / K, "steps": K, "warmup": W, "num_envs_per_gpu": N, "total_envs": total_envs, "terrain": terrain_name, "dr": bool(dr),
           "kernel": env._abi.step_kernel(), "physics_kernel_ms": phys_ms, "task_kernel_ms": task_ms,
           "state_finite": bool(torch.isfinite(state.data.qpos).all().item())}
    env.close()
    return rec


def train_step_record(torch, dev, cfg, local_rank, rank, world, table, barrier, max_over_ranks, n=4096):
    """One full PPO training step at the reference hyper-parameters (training/train.py:135-161: 2 unrolls of 20 steps, 4 x 32
    minibatch updates), fp32 ("highest", train.py:93-94), per rank `n` envs; gradients all-reduced when world > 1. The batch size grows
    with the world (256 per rank: the reference's 256 at 4096 envs, 2048 at 8 x 4096), so every rank keeps the reference's per-GPU work
    - 2 unrolls and 5120-transition minibatches - and the record is the weak-scaling series of the training step."""
    from phase_guided_terrain_traversal_b200 import ppo
    pc = ppo.PPOConfig(num_envs=n * world, batch_size=256 * world, matmul_precision="highest")
    env, wenv, state = make_env(None, cfg, local_rank, n, rank * n, "stairs", table, 1, seed=5)
    tr = ppo.PPOTrainer(wenv, state, pc)
    for _ in range(2):
        tr.training_step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    reps = 3
    steps0 = tr.env_steps
    e0.record(torch.cuda.current_stream(dev))
    for _ in range(reps):
        m = tr.training_step()
    e1.record(torch.cuda.current_stream(dev))
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / reps
    env_steps = (tr.env_steps - steps0) // reps              # the trainer's own count: unrolls x unroll_length x envs of all ranks
    rec = {"value": env_steps / (ms * 1e-3), "unit": UNIT, "ms_per_training_step": ms, "env_steps_per_training_step": env_steps, "num_envs_per_gpu": n, "batch_size": pc.batch_size, "unrolls_per_training_step": tr.unrolls_per_step,
           "precision": "fp32 (matmul_precision highest, as training/train.py:93-94)", "learner": getattr(tr, "learner_kind", "torch autograd over library GEMMs + hand-written GAE / loss-head / clip+Adam kernels"),
           "total_loss": float(m["total_loss"])}
    env.close()
    return rec


# ------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    from phase_guided_terrain_traversal_b200 import terrain
    from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
    cfg = training_overrides(default_config())
    rank = int(os.environ.get("RANK", "0"))
    if args.terrain == "curriculum":
        args.curriculum = True
        args.terrain = f"level{(rank % 10) + 1:02d}"
    table = terrain.load_terrain(args.terrain) if args.task == "stairs" else None
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, table, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    N, offset = args.num_envs, rank * args.num_envs
    if args.total_envs:
        from phase_guided_terrain_traversal_b200 import sharding
        offset, stop = sharding.shard(args.total_envs, rank, world)
        N = stop - offset
    env, wenv, state = make_env(args, cfg, local_rank, N, offset, args.task, table, args.dr)
    abi = env._abi
    pool = action_pool(torch, dev, N, 1234 + rank)
    host_pool = [a.cpu().pin_memory() for a in pool]
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    W, K = max(args.warmup, 3), args.steps
    for i in range(W):
        wenv.step(state, pool[i % 16])
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    timer = Timer(torch, dev, abi, pool, barrier, max_over_ranks)

    # ---- region A: back-to-back (L2-resident state) ------------------------------------------------
    ms_resident = timer.resident(K)

    # ---- region B: one event triple per step, L2 flushed between steps (the reported `value`) ----------
    l0 = abi.launch_count()
    ms_cold, physics_ms, task_ms = timer.flushed(K)
    launches = abi.launch_count() - l0

    # ---- region C: end to end through the public API with host actions ---------------------------------
    rew_host = torch.empty((2, N), dtype=torch.float32).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(3):
        st = wenv.step(state, host_pool[i % 16])
    barrier()
    episodes_ended = 0.0
    e0.record(stream)
    for i in range(K):
        st = wenv.step(state, host_pool[i % 16])                 # H2D (pinned, async) + physics and task kernels
        rew_host[0].copy_(st.reward, non_blocking=True)          # D2H of the step's result
        rew_host[1].copy_(st.done, non_blocking=True)
        stream.synchronize()                                     # the host consumes reward/done every step
        episodes_ended += float(rew_host[1].sum())
    e1.record(stream)
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop()

    # ---- region D: the PPO rollout collector: policy MLP (tcgen05) -> physics -> task kernel (writes the
    # transition slot), unroll_length 20, everything on the device, one native call per unroll -------------------------------
    from phase_guided_terrain_traversal_b200.policy import PolicyNet
    from phase_guided_terrain_traversal_b200.rollout import RolloutCollector
    T = 20
    net = PolicyNet(device=local_rank).init_random(1 + rank)
    col = RolloutCollector(wenv, net, unroll_length=T, seed=7 + rank)
    n_unroll = max(2, min(K, 200) // T)
    for _ in range(2):
        col.collect()
    barrier()
    l0r = abi.launch_count() + net.launch_count()
    e0.record(stream)
    for _ in range(n_unroll):
        col.collect()
    e1.record(stream)
    barrier()
    ms_rollout = max_over_ranks(e0.elapsed_time(e1))
    rollout_launches = abi.launch_count() + net.launch_count() - l0r

    done_rate = float(state.done.float().mean().item())
    niter = float(abi.buf["solver_niter"].float().mean().item())
    finite = bool(torch.isfinite(state.data.qpos).all().item())

    # ---- sub-records: BASELINE config[2], the strong-scaling series (config[3]'s size) and a full PPO training step ------
    extras = {}
    if not args.no_extras and not args.total_envs:
        Ks, Ws = max(20, min(K // 5, 100)), max(10, min(W, 30))
        extras["config2"] = sub_record(torch, dev, args, cfg, local_rank, rank, world, 8192, rank * 8192, "level07", 1, Ks, Ws, barrier, max_over_ranks, 8192 * world)
        from phase_guided_terrain_traversal_b200 import sharding
        so, se = sharding.shard(args.strong_total, rank, world)
        extras["strong"] = sub_record(torch, dev, args, cfg, local_rank, rank, world, se - so, so, "level1", 0, Ks, Ws, barrier, max_over_ranks, args.strong_total)
        extras["strong"]["note"] = f"{args.strong_total} envs in total split by index over {world} GPU(s): value(n_gpus = N) / value(n_gpus = 1) is the strong-scaling ratio"
        try:
            extras["train_step"] = train_step_record(torch, dev, cfg, local_rank, rank, world, table if args.task == "stairs" else terrain.load_terrain("level1"), barrier, max_over_ranks)
        except Exception as ex:    # the learner is a "next" row: its failure must not take the step benchmark down
            extras["train_step"] = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        peak, peak_src = load_peaks()
        total_envs = args.total_envs if args.total_envs else N * world
        value = total_envs * K / (ms_cold * 1e-3)
        clocks = sampler.summary()
        sm_mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
        fp32_peak = N_SM * FP32_LANES * 2 * sm_mhz * 1e6 / 1e12
        kname = abi.step_kernel()
        mkey = f"{kname}|{args.task}|{args.terrain}|{N}|dr{int(bool(args.dr))}"
        rec = load_kernel_metrics().get(mkey)
        achieved = B_ALG * N / ((physics_ms + task_ms) * 1e-3) / 1e9
        roof = {"bound": "fp32-issue/latency", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "hbm_frac": achieved / peak,
                "peak_source": peak_src, "kernel": kname, "kernel_ms": physics_ms, "task_kernel": "pgtt_task_kernel<OP_TASK>" if task_ms > 0 else "(fused into the step kernel)", "task_kernel_ms": task_ms,
                "algorithmic_bytes_per_env_step": B_ALG, "fp32_peak_tflops": fp32_peak,
                "fp32_peak_source": f"{N_SM} SMs x {FP32_LANES} fp32 lanes x 2 x {sm_mhz:.0f} MHz (median SM clock of this run)",
                "traffic": None, "fp32_tflops": None, "fp32_frac": None, "metrics_key": mkey,
                "note": "the step is fp32-issue / latency bound (~170 flop per algorithmic byte; SURVEY 8d, DESIGN.md 3): `frac` (HBM) is reported because the "
                        "contract asks for it, `fp32_frac` (measured flops of the step kernel / its live duration / fp32 peak) is the figure that describes it"}
        if rec:
            roof.update({"traffic": rec["dram_bytes_read"] + rec["dram_bytes_write"], "flops_per_launch": rec["fp32_flops"],
                         "fp32_tflops": rec["fp32_flops"] / (physics_ms * 1e-3) / 1e12, "fp32_frac": rec["fp32_flops"] / (physics_ms * 1e-3) / 1e12 / fp32_peak,
                         "metrics_file": rec["metrics_file"], "ncu_warps_active_pct": rec.get("warps_active_pct"), "ncu_issue_active_pct": rec.get("issue_active_pct")})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_cold / K,
            "higher_is_better": True, "scaling": "strong" if args.total_envs else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, N, world),
            "value_l2_resident": total_envs * K / (ms_resident * 1e-3), "ms_per_step_l2_resident": ms_resident / K,
            "e2e": {"value": total_envs * K / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": N * 12 * 4, "d2h_bytes_per_step": N * 2 * 4,
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": int(launches), "launches_per_step": abi.step_launches(),
            "rollout": {"value": total_envs * T * n_unroll / (ms_rollout * 1e-3), "unit": UNIT, "unroll_length": T, "unrolls": n_unroll,
                        "ms_per_env_step_batch": ms_rollout / (T * n_unroll), "gpu_launches": int(rollout_launches),
                        "what": "policy MLP 171-512-256-128-24 (tcgen05, bf16 operands, fp32 accumulate - narrower than the reference's fp32 'highest' "
                                "collector; random init) + wrapped env step whose task kernel writes the transition slot of the [T,N,.] buffers; "
                                "back-to-back (no L2 flush): 3 launches per control step"},
            "roofline": roof,
            "clocks": clocks,
            "health": {"done_rate_last_step": done_rate, "solver_niter_mean": niter, "state_finite": finite,
                       "auto_reset_fraction": episodes_ended / (N * K),
                       "auto_reset_note": "episodes ended (terminated or truncated, then auto-reset) per env-step over the e2e region of this rank"},
        }
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args, cfg, table, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
