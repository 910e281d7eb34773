#!/bin/bash
# 8 GPUs: the contract's launch shape on the final code (C2 weak scaling, 20 steps)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m8c_bench_c2.json; cut -c1-330 gpurun_out/r2m8c_bench_c2.json
python -c "
import json; d=json.load(open('gpurun_out/r2m8c_bench_c2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
