#!/bin/bash
# Final session of round 1: full GPU suite, smoke, bench (both arms), A3C bench (TF32 / 3xTF32, 1 / 4 stream groups, graph),
# launch list of one A3C iteration, ncu captures of the learner's kernels.  Usage: bash profiles/gpu_session_h.sh <tag>
set -u
TAG=${1:-r1h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --impl reference --steps 2000 --warmup 100 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench_ref rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
for cfg in "--tf32 --groups 1" "--tf32 --groups 4" "--groups 1" "--groups 4"; do
  timeout 300 python profiles/bench_a3c.py --graph $cfg >> $OUT/bench_a3c.jsonl 2>> $OUT/bench_a3c.err; echo "bench_a3c $cfg rc=$?" | tee -a $OUT/rc.txt
done
timeout 200 python profiles/gemm_bench.py --prec tf32 > $OUT/gemm_bench_tf32.log 2>&1
A3C="python profiles/bench_a3c.py --tf32 --iters 1 --warmup 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 560 --csv --log-file $OUT/a3c_launches.csv $A3C > $OUT/ncu_a3c_launches.log 2>&1
echo "ncu a3c launches rc=$?" | tee -a $OUT/rc.txt
bash profiles/ncu_a3c.sh $TAG | tee -a $OUT/rc.txt
cat $OUT/rc.txt; tail -3 $OUT/pytest_gpu.log; tail -1 $OUT/smoke.log; cat $OUT/bench_a3c.jsonl; cat $OUT/bench.json
