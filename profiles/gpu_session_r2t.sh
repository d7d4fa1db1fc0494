set -x
timeout 900 python -m pytest tests/test_gpu_a3c.py tests/test_gpu_gemm.py -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err; tail -c 400 gpurun_out/r2t_bench_n1.err
