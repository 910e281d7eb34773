# Session 5: the register diet of the hop loop (rate record not carried between arrival and event; square root and divisions of
# the chain walk without slow-path calls: 96 -> 88 registers at 5 blocks per SM, 76 without spills at 6, 72 with 8 bytes at 7):
# parity suite, then blocks per SM on C2 / C4 / C5.
mkdir -p gpurun_out; L=gpurun_out/r2al_ab.log; rm -f $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2al_pytest.log
run() { echo "== $*" >> $L; timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
run --occupancy 5
run --occupancy 6
run --occupancy 7
run --occupancy 5
run --occupancy 6
run --occupancy 7
run --occupancy 6 --excitons 4000000
run --occupancy 7 --excitons 4000000
run --occupancy 5 --workload C4 --steps 6
run --occupancy 6 --workload C4 --steps 6
run --occupancy 7 --workload C4 --steps 6
run --occupancy 5 --workload C5 --steps 4
run --occupancy 6 --workload C5 --steps 4
run --occupancy 7 --workload C5 --steps 4
run --occupancy 5 --workload C1
run --occupancy 6 --workload C1
cat $L
