# round 2, second GPU pass: parity suite, then A/B of the trap solver / parking / log on the bench workload
mkdir -p gpurun_out; rm -f gpurun_out/r2b_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest.log; cat gpurun_out/r2b_pytest.log
run() { echo "== $*" >> gpurun_out/r2b_ab.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2b_ab.log 2>&1; }
run --opt deep_thr=0
CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_libmlog.so run --opt deep_thr=0
run
run --opt deep_thr=8
run --opt deep_thr=12
run --opt deep_thr=16 --opt deep_blocks=2
run --opt deep_thr=16 --opt park_min_s=8 --opt park_min_e=4
run --opt deep_thr=8 --opt park_min_s=8 --opt park_min_e=4
run --opt deep_thr=16 --chunk 32
run --excitons 4000000 --opt deep_thr=0
run --excitons 4000000
run --excitons 4000000 --opt deep_thr=8
run --excitons 4000000 --opt deep_thr=16 --opt park_min_s=8 --opt park_min_e=4
cat gpurun_out/r2b_ab.log
