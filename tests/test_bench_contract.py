"""bench.py's reference arm runs on the host cores (the CPU oracle port), so its side of the driver contract can be
checked without a GPU: one JSON line, the contract keys, rank 0 only under torchrun."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
        "config", "impl", "cpu_baseline", "e2e"}


def _check(line, n_gpus):
    d = json.loads(line)
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == n_gpus and d["warmup"] >= 1 and d["value"] > 0 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1", "--num-envs", "64"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_prints_on_rank_zero_only():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29591", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1", "--num-envs", "64"],
                         cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 2)


import pytest  # noqa: E402


@pytest.mark.gpu
def test_b200_arm_line_carries_roofline_baseline_e2e_and_clocks():
    """The product arm on one GPU, short run: the keys the driver and the judge read, with sane values."""
    out = subprocess.run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3", "--cpu-seconds", "2"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert (KEYS - {"impl"}) | {"roofline", "gpu_launches", "clocks", "rollout", "config2", "strong", "train_step"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3 and d["scaling"] == "weak" and d["launches_per_step"] in (1, 2)
    assert d["gpu_launches"] == 20 * d["launches_per_step"]   # 4096 envs: one fused launch per step (PGTT_FUSE_TASK=0: physics + task kernel)
    r = d["roofline"]
    assert r["bound"] == "fp32-issue/latency" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["kernel"].startswith("pgtt_")
    assert r["hbm_frac"] == r["frac"] and 0 < r["kernel_ms"] <= d["ms_per_step"] * 1.001 and 60 < r["fp32_peak_tflops"] < 80
    if r["traffic"] is not None:     # a committed ncu record exists for this configuration: measured flops, not an estimate
        assert (ROOT / r["metrics_file"]).exists() and 0 < r["fp32_frac"] < 1 and r["traffic"] > 0 and r["flops_per_launch"] > 1e8
    assert 0 <= d["health"]["auto_reset_fraction"] < 0.2
    assert d["config2"]["num_envs_per_gpu"] == 8192 and d["config2"]["dr"] is True and d["config2"]["value"] > 0 and d["config2"]["state_finite"]
    assert d["strong"]["total_envs"] == 32768 and d["strong"]["value"] > 0
    assert "unavailable" in d["train_step"] or d["train_step"]["value"] > 0
    assert d["rollout"]["gpu_launches"] == (1 + d["launches_per_step"]) * d["rollout"]["unroll_length"] * d["rollout"]["unrolls"] + d["rollout"].get("extra_launches", 0)
    assert d["e2e"]["h2d_bytes_per_step"] == 4096 * 12 * 4 and d["e2e"]["d2h_bytes_per_step"] == 4096 * 2 * 4 and 0 < d["e2e"]["value"] <= d["value"] * 1.05
    assert d["cpu_baseline"]["kind"] == "port" and 0 < d["cpu_baseline"]["value"] < d["value"] / 10      # the north-star's >= 10x
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["health"]["state_finite"] is True
