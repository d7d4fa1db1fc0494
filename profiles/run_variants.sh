#!/bin/bash
# A/B of tuning variants: one short bench per "variant[@tile_bytes]" (variant = .so built by build_variants.py, or "lib").
TAG=${1:-var}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  v=${spec%@*}; n=""; [[ "$spec" == *@* ]] && n=${spec#*@}
  so=$PWD/drl_uav_cellularnet_b200/variants/$v.so; [ "$v" == "lib" ] && so=$PWD/drl_uav_cellularnet_b200/libuavenv.so
  UAVENV_TILE_BYTES=$n UAVENV_SO=$so python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 50 > $OUT/$spec.json 2> $OUT/$spec.err
  echo "$spec rc=$? $(python -c "import json;d=json.load(open('$OUT/$spec.json'));print(d['roofline']['launch_us'], d['roofline']['frac'], d['ue_steps_per_s'], d['launch_plan'])")"
done
