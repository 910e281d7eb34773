# Session 5: last check of the committed state -- smoke, GPU suite, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py 2>&1 | grep '^{' | tail -1 > gpurun_out/r2ba_bench_default.json; cut -c1-200 gpurun_out/r2ba_bench_default.json
