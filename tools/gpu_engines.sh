# GPU check of the two-engine hop kernel: parity suite, then the bench over engines / hot_pct / fast_rounds
mkdir -p gpurun_out; rm -f gpurun_out/engines.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { echo "== $*" >> gpurun_out/engines.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/engines.log 2>&1; }
run --engines 0 --fast-rounds 0
run --engines 0 --fast-rounds 2
E="--engines 1 --opt refill_min=12 --opt park_min=12 --opt park_max=32"
run $E
run $E --hot-pct 20
run $E --hot-pct 40
run $E --fast-rounds 16
run $E --chunk 128 --intervals 128
run --engines 0 --fast-rounds 2 --chunk 128 --intervals 128
run $E --excitons 4000000
run --engines 0 --fast-rounds 2 --excitons 4000000
cat gpurun_out/engines.log
