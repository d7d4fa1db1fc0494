#!/bin/bash
# One gpurun call at HEAD: GPU parity tests, smoke, bench (both arms), dense bench, ncu launch list + full capture of the
# step kernel, and the launch list of one A3C iteration.  Usage: bash profiles/gpu_session_f.sh <tag>
set -u
TAG=${1:-r1f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --impl reference --steps 2000 --warmup 100 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench_ref rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
timeout 300 python bench.py --workload dense --steps 300 --no-cpu-baseline > $OUT/bench_dense.json 2> $OUT/bench_dense.err; echo "bench_dense rc=$?" | tee -a $OUT/rc.txt
timeout 300 python profiles/bench_a3c.py --tf32 --graph > $OUT/bench_a3c.json 2> $OUT/bench_a3c.err; echo "bench_a3c rc=$?" | tee -a $OUT/rc.txt
SHORT="python bench.py --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline"
timeout 200 $SHORT > $OUT/plain_short.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $SHORT > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/rc.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o $OUT/prof_env_kernel $SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/rc.txt
A3C="python profiles/bench_a3c.py --tf32 --iters 1 --warmup 1"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 640 --csv --log-file $OUT/a3c_launches.csv $A3C > $OUT/ncu_a3c_launches.log 2>&1
echo "ncu a3c launches rc=$?" | tee -a $OUT/rc.txt
cat $OUT/rc.txt; tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json; cat $OUT/bench_ref.json; cat $OUT/bench_dense.json; cat $OUT/bench_a3c.json
