# Session 5: the event's record as ONE quad (top entries as 2^16-draw block intervals, three destinations, Gamma): a hop is two
# 256-bit loads and no rate field is carried in registers (72 registers without spills at 7 blocks per SM, 64 with 16 bytes at 8).
# Parity suite, then A/B against the previous commit's library (libcntmc_base.so), blocks per SM 7 and 8.
mkdir -p gpurun_out; L=gpurun_out/r2an_ab.log; rm -f $L
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r2an_pytest.log
run() { echo "== $*" >> $L; env $1 timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
B=CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_base.so
N=CNTMC_X=0
run $B
run $N
run $N --occupancy 8
run $N --occupancy 6
run $B
run $N
run $N --occupancy 8
run $B --excitons 4000000
run $N --excitons 4000000
run $N --excitons 4000000 --occupancy 8
run $B --workload C4 --steps 6
run $N --workload C4 --steps 6
run $N --workload C4 --steps 6 --occupancy 8
run $B --workload C5 --steps 4
run $N --workload C5 --steps 4
run $N --workload C5 --steps 4 --occupancy 8
run $B --workload C1
run $N --workload C1
cat $L
