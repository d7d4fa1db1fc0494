set -x
mkdir -p gpurun_out/r3p
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3p/gputests.log; tail -2 gpurun_out/r3p/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3p/bench_ref.json 2>/dev/null
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3p/bench.json 2> gpurun_out/r3p/bench.err
python profiles/sparse_bwd_bench.py 2>/dev/null | tail -1 > gpurun_out/r3p/sparse_bwd_bench.json
