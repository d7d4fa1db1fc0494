set -x
timeout 900 python -m pytest tests/test_gpu_api.py tests/test_gpu_a3c.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3n_bench.json 2>/dev/null
