set -x
timeout 1200 python -m pytest tests/test_gpu_a3c.py tests/test_gpu_parity.py -m gpu -x -q -k "evaluation_driver or sweep" 2>&1 | tail -6
