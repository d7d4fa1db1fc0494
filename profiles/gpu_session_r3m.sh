set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not sweep and not full_size" 2>&1 | tail -3
python bench.py --obs none --envs 8192 --precision fp32 --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/r3m_noobs.json 2>/dev/null
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3m_bench.json 2>/dev/null
