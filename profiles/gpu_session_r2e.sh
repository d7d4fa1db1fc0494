set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ${@:2} > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err
echo rc=$?
tail -c 1500 gpurun_out/r2e_bench_n$N.err
