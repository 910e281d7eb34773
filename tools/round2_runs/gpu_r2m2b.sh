#!/bin/bash
# 2 GPUs: the product's multi-GPU layer and the torchrun bench after the last session's changes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4 | tee gpurun_out/r2m2b_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m2b_bench_c2.json; cut -c1-400 gpurun_out/r2m2b_bench_c2.json
timeout 600 $TR bench.py --gpus 2 --workload F1 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m2b_bench_f1.json; cut -c1-300 gpurun_out/r2m2b_bench_f1.json
timeout 600 $TR bench.py --gpus 2 --impl reference --steps 2 --warmup 1 --cpu-budget 20 2>&1 | grep '^{' | tail -1 | cut -c1-200
