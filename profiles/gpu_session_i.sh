#!/bin/bash
# Profiles of the final round-1 learner: launch list of one A3C iteration, ncu sections of the no-observation env step and
# of the dense-layer kernels, GEMM shape timings, A3C bench lines.  Usage: bash profiles/gpu_session_i.sh <tag>
set -u
TAG=${1:-r1i}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for cfg in "--tf32 --groups 1" "--tf32 --groups 4" "--groups 1" "--groups 4"; do
  timeout 300 python profiles/bench_a3c.py --graph $cfg >> $OUT/bench_a3c.jsonl 2>> $OUT/bench_a3c.err; echo "bench_a3c $cfg rc=$?" | tee -a $OUT/rc.txt
done
timeout 200 python profiles/gemm_bench.py --prec tf32 > $OUT/gemm_bench_tf32.log 2>&1
timeout 200 python profiles/gemm_bench.py --prec fp32 > $OUT/gemm_bench_fp32.log 2>&1
timeout 200 python bench.py --obs none --envs 8192 --steps 500 --no-cpu-baseline > $OUT/bench_noobs.json 2>/dev/null
A3C="python profiles/bench_a3c.py --tf32 --iters 1 --warmup 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $OUT/a3c_launches.csv $A3C > $OUT/ncu_a3c_launches.log 2>&1
echo "ncu a3c launches rc=$?" | tee -a $OUT/rc.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:env_kernel -s 6 -c 1 -o $OUT/prof_env_small $A3C > $OUT/ncu_env.log 2>&1
echo "ncu env rc=$?" | tee -a $OUT/rc.txt
timeout 400 ncu --set full --import-source on --clock-control none -k regex:gemm_kernel -s 36 -c 10 -o $OUT/prof_gemm $A3C > $OUT/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" | tee -a $OUT/rc.txt
cat $OUT/rc.txt; cat $OUT/bench_a3c.jsonl
