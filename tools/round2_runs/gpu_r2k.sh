mkdir -p gpurun_out; rm -f gpurun_out/r2k_*.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -4 > gpurun_out/r2k_pytest.log; cat gpurun_out/r2k_pytest.log
for a in "1000000 16 deep_group=1" "1000000 8 deep_group=1" "1000000 16 deep_group=1 deep_rounds=1" "1000000 30 deep_group=1" "1000000 8 deep_group=1 deep_rounds=1" "4000000 16 deep_group=1"; do echo "== $a" >> gpurun_out/r2k_diag.log; timeout 200 python tools/deep_diag.py $a 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lane_ms', round(d['kernels_ms']-d['deep_ms'],3), 'trap_ms', d['deep_ms'], 'hops', d['hops'], 'class4', d['class4'], 'deferred(last)', d['deferred_last_round'], 'returned', d['returned'])" >> gpurun_out/r2k_diag.log 2>&1; done
cat gpurun_out/r2k_diag.log
