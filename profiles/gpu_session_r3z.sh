# A/B on one box: margin tests in pass C + cheaper local runner-up (new) against the previous build (variants/guard_old.so)
mkdir -p gpurun_out/r3z
for i in 1 2 3; do
  for v in new old; do
    if [ $v = old ]; then export UAVENV_SO=$PWD/drl_uav_cellularnet_b200/variants/guard_old.so; else unset UAVENV_SO; fi
    python bench.py --workload dense --precision fp32_guarded --no-extras --no-cpu-baseline --steps 100 --warmup 5 2>/dev/null | grep -o '^{.*' > gpurun_out/r3z/dense_guarded_${v}_$i.json
    echo $v $i $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r3z/dense_guarded_${v}_$i.json | head -1)
  done
done
unset UAVENV_SO
python bench.py --workload dense --precision fp32 --no-extras --no-cpu-baseline --steps 100 --warmup 5 2>/dev/null | grep -o '"ms_per_step": [0-9.]*' | head -1
