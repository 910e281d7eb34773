mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 1200 $TR bench.py --gpus $N --workload C5 --steps 3 --warmup 1 --e2e-steps 1 > gpurun_out/r2m${N}_bench_c5.log 2>&1; grep '^{"metric' gpurun_out/r2m${N}_bench_c5.log | tail -1 > gpurun_out/r2m${N}_bench_c5.json; cut -c1-300 gpurun_out/r2m${N}_bench_c5.json; tail -5 gpurun_out/r2m${N}_bench_c5.log | cut -c1-300
for o in 1 4 8; do timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline --opt host_slices=$o 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('slices $o', d['value'], 'e2e', d['e2e']['value'])"; done
