set -x
mkdir -p gpurun_out/r2w
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w/bench.json 2> gpurun_out/r2w/bench.err
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2w/bench_ref.json 2> gpurun_out/r2w/bench_ref.err
python bench.py --gpus 1 --steps 2000 --warmup 20 --no-extras --no-cpu-baseline > gpurun_out/r2w/bench_2000.json 2>/dev/null
SHORT="python bench.py --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --spinup-ms 0"
$SHORT > gpurun_out/r2w/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w/launches.csv $SHORT > gpurun_out/r2w/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/r2w/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o gpurun_out/r2w/prof_env_kernel $SHORT > gpurun_out/r2w/ncu_full.log 2>&1
echo "ncu full rc=$?"
