set -x
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err
tail -c 600 gpurun_out/r2b_bench_n1.err
