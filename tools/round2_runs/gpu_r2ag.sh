#!/bin/bash
# r2ag: quiet steps (the hop kernel's twin for time steps much shorter than a segment's flight time): parity, then BASELINE config 1
mkdir -p gpurun_out; T=r2ag
timeout 900 python -m pytest tests/test_gpu_c1.py tests/test_gpu_parity.py tests/test_gpu_statistics.py -q -x 2>&1 | tail -4
run() { echo "== $*"; timeout 400 python bench.py --workload C1 --steps 5 --warmup 3 --e2e-steps 2 --no-cpu-baseline $* 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'])"; }
( run --opt quiet_steps=0; run; run --opt quiet_steps=4; run --opt quiet_steps=8; run --opt quiet_steps=32; run --opt quiet_steps=64; run --opt quiet_steps=0 --excitons 1000000 --intervals 200 --chunk 64; run --excitons 1000000 --intervals 200 --chunk 64; run --opt quiet_steps=32 --excitons 1000000 --intervals 200 --chunk 64 ) 2>&1 | tee gpurun_out/${T}_c1.log
bash tools/gpu_ab.sh "--opt quiet_steps=-1" ; mv gpurun_out/ab.log gpurun_out/${T}_c2.log
