mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m${N}_bench_c2.json; cut -c1-200 gpurun_out/r2m${N}_bench_c2.json
timeout 1200 $TR bench.py --gpus $N --workload C3 --steps 3 --warmup 1 --e2e-steps 1 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m${N}_bench_c3.json; cut -c1-200 gpurun_out/r2m${N}_bench_c3.json
timeout 1200 $TR bench.py --gpus $N --workload C5 --steps 3 --warmup 1 --e2e-steps 1 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m${N}_bench_c5.json; cut -c1-200 gpurun_out/r2m${N}_bench_c5.json
timeout 600 python tools/run_multi.py $N 1000000 10 2>&1 | tail -1 > gpurun_out/r2m${N}_multi.log; cat gpurun_out/r2m${N}_multi.log
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
