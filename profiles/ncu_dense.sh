#!/bin/bash
# ncu --set full capture of the dense step kernel (config[3]).  Usage: bash profiles/ncu_dense.sh <tag> [obs]
TAG=${1:-ncu_dense}; OBS=${2:-f32}
OUT=gpurun_out/$TAG; mkdir -p $OUT
SHORT="python bench.py --workload dense --obs $OBS --steps 6 --warmup 3 --e2e-steps 2 --no-cpu-baseline"
$SHORT > $OUT/plain_short.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 5 -c 1 -o $OUT/prof_dense $SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
