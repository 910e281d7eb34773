# Round 2, draw-interval top entries: parity suite on the new library, then A/B against the library of the previous commit
# (cnt_film_monte_carlo_b200/libcntmc_base.so, built from `git archive <commit>`), bench workload C2 and C4.
mkdir -p gpurun_out; L=gpurun_out/r2ai_ab.log; rm -f $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2ai_pytest.log
run() { echo "== $*" >> $L; env $1 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
B=CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_base.so
N=CNTMC_X=0
run $B
run $N
run $B
run $N
run $N --occupancy 6
run $N --occupancy 4
run $B --excitons 4000000
run $N --excitons 4000000
run $B --workload C4
run $N --workload C4
run $B --workload C5
run $N --workload C5
run $B --workload C1
run $N --workload C1
cat $L
