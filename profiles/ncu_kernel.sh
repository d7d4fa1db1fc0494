#!/bin/bash
# ncu --set full capture of the step kernel on a short bench run.  Usage: bash profiles/ncu_kernel.sh <tag> [so]
TAG=${1:-ncu}; SO=${2:-}
OUT=gpurun_out/$TAG; mkdir -p $OUT
[ -n "$SO" ] && export UAVENV_SO=$PWD/$SO
SHORT="python bench.py --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline"
$SHORT > $OUT/plain_short.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o $OUT/prof_env_kernel $SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
