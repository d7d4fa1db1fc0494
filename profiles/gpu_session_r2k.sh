set -x
for t in 0 40 60 70 80 90 100; do
UAVENV_STREAM_AFTER=$t python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2k_dense_after$t.json 2>/dev/null
done
