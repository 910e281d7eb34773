# round 2, first GPU pass: parity suite, then A/B of the group solver / parking options on the bench workload
mkdir -p gpurun_out; rm -f gpurun_out/r2a_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_pytest.log
run() { echo "== $*" >> gpurun_out/r2a_ab.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2a_ab.log 2>&1; }
run --opt deep_thr=0
run
run --opt deep_thr=8
run --opt deep_thr=8 --opt deep_pct=30
run --opt deep_thr=16 --opt park_min_s=8 --opt park_min_e=4
run --opt deep_thr=8 --opt park_min_s=8 --opt park_min_e=4
run --opt deep_thr=8 --opt park_min_s=12 --opt park_min_e=6 --opt park_age=8
run --opt deep_thr=0 --opt park_min_s=8 --opt park_min_e=4
run --excitons 4000000 --opt deep_thr=0
run --excitons 4000000
run --excitons 4000000 --opt deep_thr=8 --opt park_min_s=8 --opt park_min_e=4
cat gpurun_out/r2a_ab.log
