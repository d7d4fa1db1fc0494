#!/bin/bash
# ncu --set full of the no-observation env step (8192 envs, 64-thread CTAs): the A3C rollout's env kernel
OUT=gpurun_out/${1:-ncu_env_small}; mkdir -p $OUT
CMD="python bench.py --obs none --envs 8192 --precision fp32 --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --spinup-ms 0"
$CMD > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 1 -o $OUT/prof_env_small $CMD > $OUT/ncu.log 2>&1
echo "ncu rc=$?"
