mkdir -p gpurun_out; rm -f gpurun_out/r2d_diag.log
for a in "1000000 16" "1000000 16 park_min_e=4" "1000000 16 park_min_e=8 park_age=8" "1000000 16 park_min_s=8 park_min_e=4" "1000000 16 park_min_e=3 park_age=16" "1000000 8 park_min_e=4 park_age=8" "1000000 8 hot_pct=10" "1000000 8 hot_pct=10 park_min_e=4 park_age=8" "1000000 16 occupancy=4" "1000000 16 occupancy=6"; do echo "== $a" >> gpurun_out/r2d_diag.log; timeout 200 python tools/deep_diag.py $a 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lane_ms', round(d['kernels_ms']-d['deep_ms'],3), 'deep_ms', d['deep_ms'], 'hops', d['hops'], 'class4', d['class4'], 'deferred', d['deferred'])" >> gpurun_out/r2d_diag.log 2>&1; done
cat gpurun_out/r2d_diag.log
