set -x
timeout 900 python -m pytest tests/test_gpu_a3c.py -m gpu -x -q -k "sparse_bwd or gradients or update or oracle" 2>&1 | tail -4
python profiles/sparse_bwd_bench.py 2>&1 | tail -1
