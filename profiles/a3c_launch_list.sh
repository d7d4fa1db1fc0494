#!/bin/bash
# ncu launch list (gpu__time_duration) of A3C iterations.  Usage: bash profiles/a3c_launch_list.sh <tag> [bench_a3c flags]
TAG=${1:-a3c}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
A3C="python profiles/bench_a3c.py --iters 1 --warmup 1 $*"
timeout 300 $A3C > $OUT/plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/a3c_launches.csv $A3C > $OUT/ncu.log 2>&1
echo "ncu rc=$?"; cat $OUT/plain.log | tail -2
