#!/bin/bash
# ncu --set full of the rollout-step product (8192 x 200 x 200, all TMA): where do its ~21 us go?
OUT=gpurun_out/${1:-ncu_gemm}; mkdir -p $OUT
CMD="python profiles/gemm_bench.py --only 7 --reps 5"
$CMD > $OUT/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o $OUT/prof_gemm_small $CMD > $OUT/ncu.log 2>&1
echo "ncu rc=$?"
