#!/bin/bash
# round 2, f1 evidence: the whole GPU suite, the table-build bench with the reference's cost beside it, one full ncu capture
# of the placement kernel on input.json's table
mkdir -p gpurun_out; T=f1c
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${T}_pytest_gpu.log
timeout 900 python tools/davoody_bench.py --ref 2>&1 | tee gpurun_out/${T}_bench.log | cut -c1-500
cat > /tmp/dv_one.py <<'PY'
import sys; sys.path.insert(0, ".")
from cnt_film_monte_carlo_b200 import davoody as dv
mc = {"zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11], "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21]}
t = dv.Tube(4, 2, 10); x = dv.Transfer(t, t); x.table(*dv.table_axes(mc)); x.table(*dv.table_axes(mc))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:placement_rate_kernel -s 1 -c 1 -o gpurun_out/${T}_placement -f python /tmp/dv_one.py > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_ncu.log
