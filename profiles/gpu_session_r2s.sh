set -x
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/p2p_check.py > gpurun_out/r2s_p2p_check_n$N.json 2> gpurun_out/r2s_p2p_check_n$N.err
echo rc=$?; tail -c 600 gpurun_out/r2s_p2p_check_n$N.err
