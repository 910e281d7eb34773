# Session 5: share of hot blocks on the final kernel (7 / 8 blocks per SM), finer sweep
mkdir -p gpurun_out; L=gpurun_out/r2ar_ab.log; rm -f $L
run() { echo "== $*" >> $L; timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
for h in 22 24 26 28 30; do run --hot-pct $h; done
for h in 20 25 30; do run --hot-pct $h --excitons 4000000; done
for h in 20 25 30 40; do run --hot-pct $h --workload C4 --steps 5; done
cat $L
