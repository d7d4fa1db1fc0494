set -x
mkdir -p gpurun_out/r3y
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "guarded or edge or dense" 2>&1 | tail -4
python bench.py --workload dense --precision fp32_guarded --no-extras --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/r3y/dense_guarded.json 2>/dev/null
python bench.py --workload dense --precision fp32 --no-extras --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/r3y/dense_fp32.json 2>/dev/null
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r3y/dense_guarded.json gpurun_out/r3y/dense_fp32.json
