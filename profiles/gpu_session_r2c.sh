set -x
for g in 1e-12 1e-4 1e-3 4e-3; do python bench.py --guard-db $g --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/r2c_guard_$g.json 2>gpurun_out/r2c_guard_$g.err; done
python bench.py --precision fp32 --steps 1000 --no-extras --no-cpu-baseline > gpurun_out/r2c_fp32.json 2>/dev/null
for g in 1e-12 1e-3; do python bench.py --guard-db $g --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c_guard20_$g.json 2>/dev/null; done
python bench.py --precision fp32 --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r2c_fp32_20.json 2>/dev/null
