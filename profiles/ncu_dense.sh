#!/bin/bash
# ncu --set full capture of the dense config[3] step kernel (32 BS x 2048 UE, 1024 envs).  Usage: bash profiles/ncu_dense.sh <tag> [precision]
TAG=${1:-ncu_dense}; PREC=${2:-fp32}
OUT=gpurun_out/$TAG; mkdir -p $OUT
SHORT="python bench.py --workload dense --precision $PREC --steps 8 --warmup 3 --e2e-steps 2 --no-cpu-baseline --no-extras --spinup-ms 0"
$SHORT > $OUT/plain_$PREC.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 1 -o $OUT/prof_dense_$PREC $SHORT > $OUT/ncu_$PREC.log 2>&1
echo "ncu dense $PREC rc=$?"
