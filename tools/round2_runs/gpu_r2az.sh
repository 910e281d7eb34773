# Session 5: the final kernels on 8 GPUs of one box -- bench.py C2 (weak) under torchrun, then the product's own layer
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 300 $TR bench.py --gpus 8 --steps 20 --warmup 3 --e2e-steps 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2az_bench_8gpu_c2.json; cut -c1-260 gpurun_out/r2az_bench_8gpu_c2.json
timeout 200 python tools/run_multi.py 8 1000000 10 2>&1 | tail -1 > gpurun_out/r2az_multi8.log; cut -c1-300 gpurun_out/r2az_multi8.log
