mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2l_pytest.log; cat gpurun_out/r2l_pytest.log
for o in 1 0; do timeout 600 python bench.py --workload C4 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline --opt csr_warp=$o 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 csr_warp=$o', d['value'], 'setup_s', d['config']['setup_s'], 'table_build_s', d['config']['table_build_s'])"; done
for o in 1 0; do timeout 600 python bench.py --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline --opt csr_warp=$o 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 csr_warp=$o', d['value'], 'table_build_s', d['config']['table_build_s'])"; done
