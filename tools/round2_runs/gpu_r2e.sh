mkdir -p gpurun_out; rm -f gpurun_out/r2e_diag.log gpurun_out/r2e_ab.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2e_pytest.log; cat gpurun_out/r2e_pytest.log
for a in "1000000 16" "1000000 16 deep_rounds=1" "1000000 8" "1000000 30" "1000000 12" "4000000 16"; do echo "== $a" >> gpurun_out/r2e_diag.log; timeout 200 python tools/deep_diag.py $a 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lane_ms', round(d['kernels_ms']-d['deep_ms'],3), 'deep_ms', d['deep_ms'], 'hops', d['hops'], 'class4', d['class4'], 'deferred(last)', d['deferred_last_round'], 'returned', d['returned'], 'walk', d['walk_events'], 'hops_i', d['hops_instr'])" >> gpurun_out/r2e_diag.log 2>&1; done
cat gpurun_out/r2e_diag.log
run() { echo "== $*" >> gpurun_out/r2e_ab.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/r2e_ab.log 2>&1; }
run --opt deep_thr=0
run
run --opt deep_thr=8
run --excitons 4000000
cat gpurun_out/r2e_ab.log
