set -x
mkdir -p gpurun_out/r3u
SHORT="python bench.py --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --spinup-ms 0"
$SHORT > gpurun_out/r3u/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3u/launches.csv $SHORT > gpurun_out/r3u/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
$SHORT > gpurun_out/r3u/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o gpurun_out/r3u/prof_env_kernel $SHORT > gpurun_out/r3u/ncu_full.log 2>&1
echo "ncu full rc=$?"
