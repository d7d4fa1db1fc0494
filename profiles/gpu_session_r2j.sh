set -x
for v in nored nowait; do
UAVENV_SO=$PWD/drl_uav_cellularnet_b200/variants/$v.so python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2j_dense_$v.json 2>/dev/null
done
for t in 20480 40064 61440; do
UAVENV_TILE_BYTES=$t python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2j_dense_tile$t.json 2>/dev/null
done
