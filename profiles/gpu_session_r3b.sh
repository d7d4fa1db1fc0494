set -x
timeout 900 python -m pytest tests/test_gpu_a3c.py -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3b_bench.json 2>/dev/null
bash profiles/a3c_launch_list.sh r3b --tf32 --groups 1 > /dev/null 2>&1
