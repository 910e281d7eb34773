#!/bin/bash
# round 2, session 5: refresh of every piece of evidence for the final kernel sources (one-quad event record, 7 / 8 blocks per SM):
# captures of the hop kernel (C2, C4) and of the davoody placement kernel, the davoody table bench
set -x
mkdir -p gpurun_out; T=r2au
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/${T}_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_default.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.log 2>&1
for w in C1 C4 C5; do timeout 1200 python bench.py --workload $w --steps 5 --warmup 3 --e2e-steps 2 > gpurun_out/${T}_bench_$w.log 2>&1; done
timeout 1800 python bench.py --workload C3 --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/${T}_bench_C3.log 2>&1
timeout 900 python tools/davoody_bench.py --ref > gpurun_out/${T}_davoody_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
# (the plain kernel by its demangled template arguments: with one 100-step launch per bench step the instrumented counting pass is among the first launches)
timeout 900 $NCU --kernel-name-base demangled -k 'regex:kubo_kernel<.*bool.0, .bool.0>' -s 2 -c 1 -o gpurun_out/${T}_c2_kubo -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c2_ncu_bench.log 2>&1
timeout 900 $NCU --kernel-name-base demangled -k 'regex:kubo_kernel<.*bool.0, .bool.0>' -s 1 -c 1 -o gpurun_out/${T}_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_ncu_bench.log 2>&1
cat > /tmp/dv_one.py <<'PY'
import sys; sys.path.insert(0, ".")
from cnt_film_monte_carlo_b200 import davoody as dv
mc = {"zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11], "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21]}
t = dv.Tube(4, 2, 10); x = dv.Transfer(t, t); x.table(*dv.table_axes(mc)); x.table(*dv.table_axes(mc))
PY
timeout 600 $NCU -k regex:placement_rate_kernel -s 1 -c 1 -o gpurun_out/${T}_placement -f python /tmp/dv_one.py > gpurun_out/${T}_placement_ncu.log 2>&1
set +x
tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest_gpu.log
for f in default reference C1 C3 C4 C5; do tail -1 gpurun_out/${T}_bench_$f.log | cut -c1-300; done
