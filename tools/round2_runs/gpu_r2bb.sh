# Session 5: host slices 2 / 4 / 6 once more with ten end-to-end steps per line, alternating (is 4 still the right default?)
mkdir -p gpurun_out; L=gpurun_out/r2bb_ab.log; rm -f $L
run() { echo "== $*" >> $L; timeout 300 python bench.py --steps 3 --warmup 3 --e2e-steps 10 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['e2e']['value']/d['value'])" >> $L 2>&1; }
for r in 1 2; do for k in 2 4 6; do run --opt host_slices=$k --opt slice_share=$k; done; done
cat $L
