mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
for o in 0 1; do timeout 600 python bench.py --workload C4 --steps 4 --warmup 2 --e2e-steps 1 --no-cpu-baseline --opt ahead=$o 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C4 ahead=$o', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
for o in 0 1; do timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline --opt ahead=$o 2>&1 | grep '^{"metric' | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('C2 ahead=$o', d['value'], d['ms_per_step'], d['roofline']['frac'])"; done
