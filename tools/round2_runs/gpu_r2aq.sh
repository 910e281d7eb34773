# Session 5: (a) one GPU -- what the 16 bytes of spills of the 64-register build cost at equal occupancy, hot share re-checked on the
# final kernel; run with `gpurun --gpus 1`.
mkdir -p gpurun_out; L=gpurun_out/r2aq_ab.log; rm -f $L
run() { echo "== $*" >> $L; env $1 timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('traffic'))" >> $L 2>&1; }
N=CNTMC_X=0
run $N
run CNTMC_DBG_BLOCKS_PER_SM=7 --occupancy 8
run $N --occupancy 8
run $N --hot-pct 25
run $N --hot-pct 35
run $N --chunk 64
run $N --excitons 2000000
cat $L
