#!/bin/bash
# r2u: davoody tests + table bench after the pass-count change; launch length on the bench workload
mkdir -p gpurun_out; T=r2u
timeout 900 python -m pytest tests/test_gpu_davoody.py -q 2>&1 | tail -8 > gpurun_out/${T}_pytest.log; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python tools/davoody_bench.py --ref 2>&1 | tee gpurun_out/${T}_bench.log | cut -c1-420
bash tools/gpu_ab.sh "--chunk 64" "--chunk 100" "--chunk 50" "--chunk 100 --hot-pct 20" "--chunk 100 --hot-pct 40" "--chunk 34"
mv gpurun_out/ab.log gpurun_out/${T}_ab.log
