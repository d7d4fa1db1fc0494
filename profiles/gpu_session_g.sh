#!/bin/bash
# One gpurun call: full GPU suite, smoke, A3C bench (TF32 / 3xTF32, graph), GEMM shape timings, launch list of one A3C
# iteration, ncu capture of the dense-layer kernels.  Usage: bash profiles/gpu_session_g.sh <tag>
set -u
TAG=${1:-r1g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
timeout 300 python profiles/bench_a3c.py --tf32 --graph > $OUT/bench_a3c_tf32.json 2> $OUT/bench_a3c.err; echo "bench_a3c tf32 rc=$?" | tee -a $OUT/rc.txt
timeout 300 python profiles/bench_a3c.py --graph > $OUT/bench_a3c_fp32.json 2>> $OUT/bench_a3c.err; echo "bench_a3c fp32 rc=$?" | tee -a $OUT/rc.txt
timeout 200 python profiles/gemm_bench.py --prec tf32 > $OUT/gemm_bench_tf32.log 2>&1
timeout 200 python profiles/gemm_bench.py --prec fp32 > $OUT/gemm_bench_fp32.log 2>&1
A3C="python profiles/bench_a3c.py --tf32 --iters 1 --warmup 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 560 --csv --log-file $OUT/a3c_launches.csv $A3C > $OUT/ncu_a3c_launches.log 2>&1
echo "ncu a3c launches rc=$?" | tee -a $OUT/rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 40 -c 14 -o $OUT/prof_gemm $A3C > $OUT/ncu_gemm.log 2>&1
echo "ncu gemm rc=$?" | tee -a $OUT/rc.txt
cat $OUT/rc.txt; tail -3 $OUT/pytest_gpu.log; tail -1 $OUT/smoke.log; cat $OUT/bench_a3c_tf32.json $OUT/bench_a3c_fp32.json
