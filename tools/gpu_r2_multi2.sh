mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -4 > gpurun_out/r2m2c_pytest.log; cat gpurun_out/r2m2c_pytest.log
timeout 900 python -m pytest tests/test_gpu_contacts.py -q -k replays 2>&1 | tail -3
timeout 600 python tools/run_multi.py 2 1000000 10 2>&1 | tail -1 > gpurun_out/r2m2c_multi.log; cat gpurun_out/r2m2c_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m2c_bench_c2.json; cut -c1-300 gpurun_out/r2m2c_bench_c2.json
timeout 1200 $TR bench.py --gpus 2 --workload C3 --steps 3 --warmup 1 --e2e-steps 1 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2m2c_bench_c3.json; cut -c1-300 gpurun_out/r2m2c_bench_c3.json
