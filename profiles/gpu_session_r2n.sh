set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2n_tests.log
tail -3 gpurun_out/r2n_tests.log
for p in fp32 fp32_guarded fp64; do python bench.py --workload dense --precision $p --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2n_dense_$p.json 2>/dev/null; done
python bench.py --workload dense --precision fp32 --obs none --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2n_dense_noobs.json 2>/dev/null
UAVENV_SO=$PWD/drl_uav_cellularnet_b200/variants/minb4.so python bench.py --workload dense --precision fp32 --steps 200 --no-cpu-baseline --no-extras > gpurun_out/r2n_dense_minb4.json 2>/dev/null
for p in fp32 fp32_guarded; do python bench.py --precision $p --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/r2n_default_$p.json 2>/dev/null; done
python bench.py --obs none --envs 8192 --precision fp32 --steps 1000 --no-cpu-baseline --no-extras > gpurun_out/r2n_noobs_fp32.json 2>/dev/null
bash profiles/ncu_dense.sh r2n fp32
