mkdir -p gpurun_out; rm -f gpurun_out/r2f_ab.log
run() { echo "== $CNTMC_LIB $*" >> gpurun_out/r2f_ab.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2f_ab.log 2>&1; }
D=$PWD/cnt_film_monte_carlo_b200
run --opt deep_thr=0
CNTMC_LIB=$D/libcntmc_nopark.so run --opt deep_thr=0
CNTMC_LIB=$D/libcntmc_noparkdefer.so run --opt deep_thr=0
run --opt deep_thr=0 --hot-pct 20
run --opt deep_thr=0 --hot-pct 40
CNTMC_LIB=$D/libcntmc_r1.so run
run --opt deep_thr=0
cat gpurun_out/r2f_ab.log
