mkdir -p gpurun_out; rm -f gpurun_out/r2h_*.log
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r2h_pytest.log; cat gpurun_out/r2h_pytest.log
for o in "host_slices=4" "host_slices=2" "host_slices=1" "host_slices=8" "host_slices=3"; do timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline --opt $o 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$o', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])" >> gpurun_out/r2h_e2e.log 2>&1; done; cat gpurun_out/r2h_e2e.log
timeout 300 python bench.py --workload C1 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
