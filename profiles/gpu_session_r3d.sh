set -x
for p in 1 2 4; do UAVNET_BWD_PASSES=$p python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3d_bench_passes$p.json 2>/dev/null; done
UAVNET_SPARSE_BWD=scatter python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r3d_bench_scatter.json 2>/dev/null
bash profiles/a3c_launch_list.sh r3d --tf32 --groups 1 > /dev/null 2>&1
