# Session 5: 7 against 8 blocks per SM by population (where does the eighth block start to pay on the L2-resident film?)
mkdir -p gpurun_out; L=gpurun_out/r2ay_ab.log; rm -f $L
run() { echo "== $*" >> $L; timeout 600 python bench.py --warmup 3 --e2e-steps 1 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
for p in 2000000 3000000 6000000; do for o in 7 8; do run --steps 8 --excitons $p --occupancy $o; done; done
for o in 7 8; do run --steps 3 --workload C3 --occupancy $o; done
for o in 7 8; do run --steps 3 --workload C3 --excitons-total 12500000 --occupancy $o; done
cat $L
