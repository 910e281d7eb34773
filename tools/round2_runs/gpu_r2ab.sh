#!/bin/bash
# r2ab: the committed library against one 256-bit store per staging record (libcntmc_stg256.so), alternating
mkdir -p gpurun_out
D=$PWD/cnt_film_monte_carlo_b200
run() { echo "== ${CNTMC_LIB##*/} $*"; timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])"; }
( CNTMC_LIB=$D/libcntmc_base.so run; CNTMC_LIB=$D/libcntmc_stg256.so run; CNTMC_LIB=$D/libcntmc_base.so run; CNTMC_LIB=$D/libcntmc_stg256.so run; CNTMC_LIB=$D/libcntmc_base.so run --excitons 4000000; CNTMC_LIB=$D/libcntmc_stg256.so run --excitons 4000000 ) 2>&1 | tee gpurun_out/r2ab_ab.log
