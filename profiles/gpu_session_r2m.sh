set -x
python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q -k "not sweep and not full_size" 2>&1 | tail -8 > gpurun_out/r2m_tests.log
tail -3 gpurun_out/r2m_tests.log
for sp in 10 25 50 75 100; do UAVENV_STREAM_SPREAD=$sp python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2m_dense_spread$sp.json 2>/dev/null; done
python bench.py --workload dense --precision fp32_guarded --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2m_dense_guarded.json 2>/dev/null
