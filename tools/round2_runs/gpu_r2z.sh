#!/bin/bash
# r2z: bench.py --workload F1 (the davoody table to the bench contract), both arms
mkdir -p gpurun_out
timeout 600 python bench.py --workload F1 2>&1 | grep '^{' | tail -1 > gpurun_out/r2z_bench_F1.json; cut -c1-1500 gpurun_out/r2z_bench_F1.json
timeout 600 python bench.py --workload F1 --impl reference --steps 3 2>&1 | grep '^{' | tail -1 > gpurun_out/r2z_bench_F1_reference.json; cut -c1-700 gpurun_out/r2z_bench_F1_reference.json
timeout 300 python -m pytest tests/test_gpu_davoody.py -q 2>&1 | tail -2
