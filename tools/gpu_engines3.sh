mkdir -p gpurun_out; rm -f gpurun_out/engines3.log
run() { echo "== $*" >> gpurun_out/engines3.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/engines3.log 2>&1; }
run --engines 0 --fast-rounds 0
run --engines 0 --fast-rounds 1
run --engines 0 --fast-rounds 2
run --engines 0 --fast-rounds 0
run --engines 0 --fast-rounds 1
run --engines 0 --fast-rounds 1 --excitons 4000000
run --engines 0 --fast-rounds 0 --excitons 4000000
cat gpurun_out/engines3.log
