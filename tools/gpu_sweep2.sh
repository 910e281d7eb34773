mkdir -p gpurun_out; rm -f gpurun_out/sweep2.log
run() { echo "== $*" >> gpurun_out/sweep2.log; timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/sweep2.log 2>&1; }
for ch in 4 8 16 32; do for fr in 0 4 8; do run --chunk $ch --fast-rounds $fr --intervals 128; done; done
run --chunk 64 --fast-rounds 0 --intervals 128
run --chunk 64 --fast-rounds 4 --intervals 128
cat gpurun_out/sweep2.log
