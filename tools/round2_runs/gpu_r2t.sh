#!/bin/bash
# r2t: the trap kernel beside the lane kernel (deep_overlap=1): parity, then launch times and bench lines
mkdir -p gpurun_out; T=r2t
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_davoody.py -q -x 2>&1 | tail -15 > gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
for o in "deep_overlap=0 trap_burst=4" "deep_overlap=1 trap_burst=4" "deep_overlap=1 trap_burst=4 overlap_trap_blocks=2" "deep_overlap=1 trap_burst=1" "deep_overlap=1 trap_burst=8 hot_pct=40"; do
  echo "== thr 8 group 1 $o"; timeout 120 python tools/deep_diag.py 1e6 8 deep_group=1 deep_rounds=1 $o; done 2>&1 | tee gpurun_out/${T}_diag.log | cut -c1-300
echo "== thr 16 group 1 overlap"; timeout 120 python tools/deep_diag.py 1e6 16 deep_group=1 deep_rounds=1 deep_overlap=1 trap_burst=4 2>&1 | tee -a gpurun_out/${T}_diag.log | cut -c1-300
D="--opt deep_group=1 --opt deep_rounds=1 --opt trap_burst=4"
bash tools/gpu_ab.sh "--opt deep_thr=0" "--opt deep_thr=8 $D" "--opt deep_thr=8 $D --opt deep_overlap=1" "--opt deep_thr=8 $D --opt deep_overlap=1 --opt overlap_trap_blocks=2" "--opt deep_thr=16 $D --opt deep_overlap=1"
mv gpurun_out/ab.log gpurun_out/${T}_ab.log
