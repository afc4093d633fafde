set -x
timeout 300 python -m pytest tests/test_mlp_native.py -m gpu -q -x 2>&1 | tail -25 | cut -c1-400
