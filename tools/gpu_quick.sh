# quick GPU check of a kernel change: parity suite, bench at 1e6 excitons, C4 at full size
mkdir -p gpurun_out; rm -f gpurun_out/quick_sweep.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { echo "== $*" >> gpurun_out/quick_sweep.log; timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'])" >> gpurun_out/quick_sweep.log 2>&1; }
run
run
cat gpurun_out/quick_sweep.log
timeout 600 python tools/run_c4.py 1.0 1000000 2>&1 | tail -1
