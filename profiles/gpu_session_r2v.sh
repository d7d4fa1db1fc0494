set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_a3c.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2v_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke_ncu.log 2>&1
echo ncu rc=$?
