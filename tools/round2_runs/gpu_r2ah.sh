#!/bin/bash
# r2ah: validation of the committed state at the start of the last session: smoke, GPU tests, both bench arms
mkdir -p gpurun_out; T=r2ah
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/${T}_smoke.log
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_default.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.log 2>&1
tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest_gpu.log
for f in default reference; do tail -1 gpurun_out/${T}_bench_$f.log | cut -c1-400; done
