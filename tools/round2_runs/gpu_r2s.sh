#!/bin/bash
# r2s: davoody tests after the test fixes; event bursts (option event_burst / trap_burst) on the plain loop and on pure trapped warps
mkdir -p gpurun_out; T=r2s
timeout 900 python -m pytest tests/test_gpu_davoody.py -q 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -5 gpurun_out/${T}_pytest.log
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -3
CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_before.so bash tools/gpu_ab.sh "--opt top_entries=1"; mv gpurun_out/ab.log gpurun_out/${T}_ab_before.log
bash tools/gpu_ab.sh "--opt event_burst=1" "--opt event_burst=2" "--opt event_burst=4" "--opt event_burst=8" "--opt event_burst=64"; mv gpurun_out/ab.log gpurun_out/${T}_ab.log
for b in 1 4 16 64; do echo "== thr 8 group 1 rounds 1 trap_burst $b"; timeout 200 python tools/deep_diag.py 1e6 8 deep_group=1 deep_rounds=1 trap_burst=$b; done 2>&1 | tee gpurun_out/${T}_diag.log | cut -c1-330
for b in 1 8; do echo "== thr 8 group 1 rounds 1 trap_burst 16 event_burst $b"; timeout 200 python tools/deep_diag.py 1e6 8 deep_group=1 deep_rounds=1 trap_burst=16 event_burst=$b; done 2>&1 | tee -a gpurun_out/${T}_diag.log | cut -c1-330
