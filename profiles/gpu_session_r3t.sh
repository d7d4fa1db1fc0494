set -x
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not full_size" 2>&1 | tail -3
for p in fp32_guarded fp32; do python bench.py --precision $p --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/r3t_$p.json 2>/dev/null; done
python bench.py --workload dense --precision fp32_guarded --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r3t_dense_guarded.json 2>/dev/null
python bench.py --obs none --envs 8192 --precision fp32_guarded --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/r3t_noobs_guarded.json 2>/dev/null
