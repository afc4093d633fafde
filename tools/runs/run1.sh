set -x
python -c "import jax, mujoco" 2>&1 | tail -1
python -c "import mujoco" 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for N in 2048 4096; do PGTT_KERNEL=warp python tools/kernel_times.py stairs $N level1 100; done
for N in 4096 8192; do PGTT_KERNEL=quad python tools/kernel_times.py stairs $N level1 100; done
PGTT_KERNEL=quad python tools/kernel_times.py stairs 8192 level07 100 1
PGTT_KERNEL=quad PGTT_QUAD_WARPS=1 python tools/kernel_times.py stairs 4096 level1 100
PGTT_KERNEL=quad PGTT_QUAD_WARPS=4 python tools/kernel_times.py stairs 4096 level1 100
