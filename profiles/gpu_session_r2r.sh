set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/p2p_check.py > gpurun_out/r2r_p2p_check_n$N.json 2> gpurun_out/r2r_p2p_check_n$N.err
echo rc=$?; tail -c 600 gpurun_out/r2r_p2p_check_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 --push p2p > gpurun_out/r2r_bench_p2p_n$N.json 2> gpurun_out/r2r_bench_p2p_n$N.err
echo rc=$?; tail -c 600 gpurun_out/r2r_bench_p2p_n$N.err
