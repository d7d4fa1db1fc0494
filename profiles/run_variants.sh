#!/bin/bash
# A/B of the tuning variants built by profiles/build_variants.py: one short bench per variant.
TAG=${1:-var}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
for v in "$@"; do
  UAVENV_SO=$PWD/drl_uav_cellularnet_b200/variants/$v.so python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 50 > $OUT/$v.json 2> $OUT/$v.err
  echo "$v rc=$? $(python -c "import json;d=json.load(open('$OUT/$v.json'));print(d['roofline']['launch_us'], d['roofline']['frac'], d['ue_steps_per_s'])")"
done
