#!/bin/bash
# ncu --set full capture of the learner's kernels (one launch each) on a short A3C run.  Usage: bash profiles/ncu_a3c.sh <tag>
TAG=${1:-ncu_a3c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
SHORT="python profiles/bench_a3c.py --tf32 --iters 2 --warmup 1"
$SHORT > $OUT/plain_short.log 2>&1 && \
ncu --section SpeedOfLight --section LaunchStats --section Occupancy --section MemoryWorkloadAnalysis --clock-control none -k regex:'sparse_fwd_kernel|sparse_bwd_kernel|actor_head_bwd_kernel|softmax_sample_kernel|rmsprop_kernel|env_kernel' -s 40 -c 46 -o $OUT/prof_a3c $SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?"
