# dense per-(UE, 4 BS) loop: two UE groups per lane in flight (UAVENV_QUAD_UNROLL=2) at 3 and at 2 CTAs per SM, against the default
mkdir -p gpurun_out/r3ae
for v in default u2m3 u2m2 default u2m3 u2m2; do
  if [ $v = default ]; then unset UAVENV_SO; else export UAVENV_SO=$PWD/drl_uav_cellularnet_b200/variants/$v.so; fi
  for p in fp32_guarded fp32; do
    echo $v $p $(python bench.py --workload dense --precision $p --no-extras --no-cpu-baseline --steps 60 --warmup 5 --e2e-steps 2 2>/dev/null | grep -o '^{.*' | grep -o '"ms_per_step": [0-9.]*' | head -1) | tee -a gpurun_out/r3ae/results.txt
  done
done
