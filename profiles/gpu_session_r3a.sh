set -x
timeout 900 python -m pytest tests/test_gpu_a3c.py tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3a_bench.json 2>/dev/null
