# Session 5: the logarithm's polynomial coefficients from the constant bank (LDCU.128 pairs) instead of immediates (two UMOV each):
# A/B against the committed library, alternating
mkdir -p gpurun_out; L=gpurun_out/r2at_ab.log; rm -f $L
run() { echo "== $*" >> $L; env $1 timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
V=CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_lc.so
N=CNTMC_X=0
run $N
run $V
run $N
run $V
run $N --workload C4 --steps 5
run $V --workload C4 --steps 5
run $N --workload C5 --steps 4
run $V --workload C5 --steps 4
cat $L
