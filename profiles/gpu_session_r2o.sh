set -x
for t in 1 2 4 7 12; do UAVENV_STREAM_TURNS=$t python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2o_dense_turns$t.json 2>/dev/null; done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not sweep and not full_size" 2>&1 | tail -3
