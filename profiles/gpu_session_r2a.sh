set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "guarded or fp32_full or sweep or full_size" 2>&1 | tail -40 > gpurun_out/r2a_parity.log
for p in fp32 fp32_guarded fp64; do python bench.py --precision $p --steps 1000 --no-cpu-baseline > gpurun_out/r2a_bench_$p.json 2> gpurun_out/r2a_bench_$p.err; done
for p in fp32 fp32_guarded; do python bench.py --workload dense --precision $p --steps 200 --no-cpu-baseline > gpurun_out/r2a_dense_$p.json 2> gpurun_out/r2a_dense_$p.err; done
for p in fp32 fp32_guarded; do python bench.py --obs none --envs 8192 --precision $p --steps 1000 --no-cpu-baseline > gpurun_out/r2a_noobs_$p.json 2> gpurun_out/r2a_noobs_$p.err; done
tail -3 gpurun_out/r2a_parity.log
