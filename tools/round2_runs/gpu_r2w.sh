#!/bin/bash
# r2w: the next event's Philox block computed behind the destination load (libcntmc_ahead.so) against the committed library; then
# the hand-over under initcheck
mkdir -p gpurun_out
D=$PWD/cnt_film_monte_carlo_b200
run() { echo "== ${CNTMC_LIB##*/} $*"; timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])"; }
CNTMC_LIB=$D/libcntmc_ahead.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "replay_of_reference or philox_matches or chunking" 2>&1 | tail -2
( run; CNTMC_LIB=$D/libcntmc_ahead.so run; run; CNTMC_LIB=$D/libcntmc_ahead.so run; run --excitons 4000000; CNTMC_LIB=$D/libcntmc_ahead.so run --excitons 4000000; run --workload C4 --steps 3; CNTMC_LIB=$D/libcntmc_ahead.so run --workload C4 --steps 3 ) 2>&1 | tee gpurun_out/r2w_ab.log
sed -n '/^# the in-launch/,$p' tools/gpu_r2_sanitize2.sh > /tmp/ov.sh; rm -f gpurun_out/r2b_sanitizer_summary.log; bash /tmp/ov.sh
