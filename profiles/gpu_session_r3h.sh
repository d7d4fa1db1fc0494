set -x
mkdir -p gpurun_out/r3h
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3h/gputests.log; tail -2 gpurun_out/r3h/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3h/bench_ref.json 2>/dev/null
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3h/bench.json 2> gpurun_out/r3h/bench.err
