mkdir -p gpurun_out; rm -f gpurun_out/check3.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*" >> gpurun_out/check3.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/check3.log 2>&1; }
run
run --top-entries 0
run
cat gpurun_out/check3.log
echo "== C5 share, top entries on / off"
timeout 300 python tools/run_c5.py 2000000 100 25 2>&1 | tail -2
CNTMC_C5_TOP=0 timeout 300 python tools/run_c5.py 2000000 100 25 2>&1 | tail -2
echo "== C4"
timeout 600 python tools/run_c4.py 1.0 1000000 2>&1 | tail -1
