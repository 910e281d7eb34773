# GPU check of the fast event path: parity suite, then the bench at several fast_rounds settings
mkdir -p gpurun_out; rm -f gpurun_out/fast_sweep.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "== $*" >> gpurun_out/fast_sweep.log; timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/fast_sweep.log 2>&1; }
for fr in 0 8 2 4 16 32; do run --fast-rounds $fr; done
run --fast-rounds 8 --hot-pct 20
run --fast-rounds 8 --hot-pct 50
run --fast-rounds 8 --excitons 4000000
run --fast-rounds 0 --excitons 4000000
cat gpurun_out/fast_sweep.log
