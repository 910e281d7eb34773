# Session 5: two more registers off the loop (Philox block validity as a flag instead of a block number, re-injections counted at
# once): the 64-register build's spills shrink from 62 to 16 bytes, all in the cold off-site normalise.  Parity, then 7 against 8 blocks.
mkdir -p gpurun_out; L=gpurun_out/r2as_ab.log; rm -f $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2as_pytest.log
run() { echo "== $*" >> $L; timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
run --occupancy 7
run --occupancy 8
run --occupancy 7
run --occupancy 8
run --occupancy 8 --hot-pct 30
run --occupancy 8 --hot-pct 20
run --occupancy 7 --excitons 4000000
run --occupancy 8 --excitons 4000000
run --occupancy 7 --workload C4 --steps 5
run --occupancy 8 --workload C4 --steps 5
run --occupancy 7 --workload C5 --steps 4
run --occupancy 8 --workload C5 --steps 4
cat $L
