#!/bin/bash
# r2aa: one 256-bit store per staging record (always on) and the segment-time windows (option seg_windows) on C2 / C4 / C5
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_c1.py -q -x 2>&1 | tail -3
bash tools/gpu_ab.sh "--opt seg_windows=0" "--opt seg_windows=1" "--opt seg_windows=0" "--opt seg_windows=1" "--excitons 4000000 --opt seg_windows=0" "--excitons 4000000 --opt seg_windows=1"
mv gpurun_out/ab.log gpurun_out/r2aa_ab.log
run() { echo "== $*"; timeout 400 python bench.py --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])"; }
( run --workload C4 --opt seg_windows=0; run --workload C4 --opt seg_windows=1; run --workload C5 --opt seg_windows=0; run --workload C5 --opt seg_windows=1 ) 2>&1 | tee gpurun_out/r2aa_c45.log
