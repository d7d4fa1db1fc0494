#!/bin/bash
# ncu capture of the learner's own kernels (one launch each, selected sections) on a short A3C run.
# Usage: bash profiles/ncu_a3c.sh <tag>
TAG=${1:-ncu_a3c}
OUT=gpurun_out/$TAG; mkdir -p $OUT
SHORT="python profiles/bench_a3c.py --tf32 --iters 1 --warmup 1"
timeout 300 $SHORT > $OUT/plain_short.log 2>&1 && \
for K in sparse_fwd_kernel sparse_bwd_kernel actor_head_bwd_kernel softmax_sample_kernel rmsprop_kernel rank1_mask_kernel nstep_targets_kernel; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$K -s 1 -c 1 -o $OUT/prof_$K $SHORT > $OUT/ncu_$K.log 2>&1
  echo "ncu $K rc=$?"
done
