set -x
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_a3c.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3g_bench.json 2>/dev/null
bash profiles/ncu_gemm_small.sh r3g
