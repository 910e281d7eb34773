#!/bin/bash
# r2v: resident warps on the HBM-gather workload (C4) and on contacts (C5)
mkdir -p gpurun_out; T=r2v; rm -f gpurun_out/${T}.log
run() { echo "== $*" >> gpurun_out/${T}.log; timeout 400 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/${T}.log 2>&1; }
run --workload C4 --occupancy 5
run --workload C4 --occupancy 6
run --workload C4 --occupancy 4
run --workload C4 --occupancy 6 --hot-pct 50
run --workload C4 --hot-pct 50
run --workload C4 --excitons 4000000
cat gpurun_out/${T}.log
