#!/bin/bash
# re-stamp the hash-stamped ncu counters after the last edit of kernels.cuh (the default hop kernel's code is unchanged)
mkdir -p gpurun_out; T=r2ae
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:kubo_kernel -s 3 -c 1 -o gpurun_out/${T}_c2_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c2_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/${T}_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_ncu_bench.log 2>&1
grep -h '^{"metric' gpurun_out/${T}_c2_ncu_bench.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c2 hops/step', d['config']['hops_per_step'])"
grep -h '^{"metric' gpurun_out/${T}_c4_ncu_bench.log | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('c4 hops/step', d['config']['hops_per_step'])"
timeout 600 python bench.py 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/${T}_bench_default.json
