set -x
mkdir -p gpurun_out/r3w
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3w/gputests.log; tail -2 gpurun_out/r3w/gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/r3w/bench_ref.json 2>/dev/null
python bench.py > gpurun_out/r3w/bench.json 2> gpurun_out/r3w/bench.err
