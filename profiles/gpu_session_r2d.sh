set -x
mkdir -p gpurun_out/r2d
for p in fp32_guarded fp32; do
  python bench.py --precision $p --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2d/spin_$p.json 2>/dev/null
  python bench.py --precision $p --steps 20 --warmup 5 --no-extras --no-cpu-baseline --spinup-ms 0 > gpurun_out/r2d/nospin_$p.json 2>/dev/null
done
for p in fp32_guarded fp32; do
  SHORT="python bench.py --precision $p --steps 12 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --spinup-ms 0"
  $SHORT > gpurun_out/r2d/plain_$p.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 6 -c 2 -o gpurun_out/r2d/prof_$p $SHORT > gpurun_out/r2d/ncu_$p.log 2>&1
  echo "ncu $p rc=$?"
done
