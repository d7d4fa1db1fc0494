set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/p2p_check.py --split --iters 10 2>/dev/null | grep "^{" > gpurun_out/r3r_p2p_check_split_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r3r_bench_split_n$N.json 2> gpurun_out/r3r_bench_split_n$N.err
echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 bench.py --gpus $N --steps 20 --warmup 5 --no-split-push > gpurun_out/r3r_bench_nosplit_n$N.json 2> gpurun_out/r3r_bench_nosplit_n$N.err
echo rc=$?
