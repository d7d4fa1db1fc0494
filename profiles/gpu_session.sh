#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list, ncu full capture of the step kernel.
# Usage (from the repo root on the GPU box): bash profiles/gpu_session.sh <tag>
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/rc.txt
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/rc.txt
python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" | tee -a $OUT/rc.txt
python bench.py --impl reference --steps 2000 --warmup 100 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench_ref rc=$?" | tee -a $OUT/rc.txt
python profiles/write_ceiling.py > $OUT/write_ceiling.json 2> $OUT/write_ceiling.err; echo "write_ceiling rc=$?" | tee -a $OUT/rc.txt
SHORT="python bench.py --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline"
$SHORT > $OUT/plain_short.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches.csv $SHORT > $OUT/ncu_launches.log 2>&1
echo "ncu launches rc=$?" | tee -a $OUT/rc.txt
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o $OUT/prof_env_kernel $SHORT > $OUT/ncu_full.log 2>&1
echo "ncu full rc=$?" | tee -a $OUT/rc.txt
cat $OUT/rc.txt; tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/write_ceiling.json; cat $OUT/bench.json; cat $OUT/bench_ref.json
