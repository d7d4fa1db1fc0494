set -x
for g in 2 8 16; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --a3c-groups $g > gpurun_out/r2x_groups$g.json 2>/dev/null; done
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --a3c-groups 4 --a3c-precision fp32 > gpurun_out/r2x_groups4_fp32.json 2>/dev/null
