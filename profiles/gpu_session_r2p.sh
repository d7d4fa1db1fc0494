set -x
for p in fp32 fp32_guarded; do python bench.py --workload dense --precision $p --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2p_dense_$p.json 2>/dev/null; done
python bench.py --workload dense --precision fp32 --obs none --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2p_dense_noobs.json 2>/dev/null
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not sweep and not full_size" 2>&1 | tail -3
