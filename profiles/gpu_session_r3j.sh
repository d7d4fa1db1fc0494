set -x
timeout 600 python -m pytest tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do python bench.py --steps 200 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r3j_bench_$i.json 2>/dev/null; done
