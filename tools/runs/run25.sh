set -x
for i in 1 2; do PGTT_KERNEL=warp python tools/kernel_times.py stairs 4096 level1 200 2>&1 | grep back; done
timeout 900 python -m pytest tests/test_kernel_parity.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-300
