# Round 2: C4 (rows in HBM) against resident blocks per SM and the share of hot blocks -- does the HBM-latency regime want more warps?
mkdir -p gpurun_out; L=gpurun_out/r2aj_c4sweep.log; rm -f $L
run() { echo "== $*" >> $L; timeout 400 python bench.py --workload C4 --steps 6 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
run
run --occupancy 4
run --occupancy 6
run --occupancy 8
run --hot-pct 50
run --hot-pct 15
cat $L
