set -x
N=${1:-8}
for v in "0 64" "0 148" "0 592"; do
set -- $v
UAVNET_P2P_DBG=$1 UAVNET_P2P_GRID=$2 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 profiles/p2p_check.py --iters 20 2>/dev/null | grep "^{" > gpurun_out/r2u_p2p_dbg$1_grid$2.json
done
