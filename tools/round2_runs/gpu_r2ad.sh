#!/bin/bash
# r2ad: the trap kernel beside the lane kernel again, after taking the acquire load (and its L1 invalidation) out of the empty polls
mkdir -p gpurun_out; T=r2ad
timeout 300 python tools/overlap_check.py 2>&1 | tail -1
for o in "deep_overlap=0 trap_burst=4" "deep_overlap=1 trap_burst=4" "deep_overlap=1 trap_burst=4 overlap_trap_blocks=2" "deep_overlap=1 trap_burst=4 overlap_trap_blocks=3"; do
  echo "== thr 8 group 1 $o"; timeout 120 python tools/deep_diag.py 1e6 8 deep_group=1 deep_rounds=1 $o; done 2>&1 | tee gpurun_out/${T}_diag.log | cut -c1-240
echo "== thr 16 overlap 2"; timeout 120 python tools/deep_diag.py 1e6 16 deep_group=1 deep_rounds=1 deep_overlap=1 trap_burst=4 overlap_trap_blocks=2 2>&1 | tee -a gpurun_out/${T}_diag.log | cut -c1-240
echo "== thr 30 overlap 1"; timeout 120 python tools/deep_diag.py 1e6 30 deep_group=1 deep_rounds=1 deep_overlap=1 trap_burst=4 overlap_trap_blocks=1 2>&1 | tee -a gpurun_out/${T}_diag.log | cut -c1-240
D="--opt deep_group=1 --opt deep_rounds=1 --opt trap_burst=4"
bash tools/gpu_ab.sh "--opt deep_thr=0" "--opt deep_thr=8 $D --opt deep_overlap=1" "--opt deep_thr=8 $D --opt deep_overlap=1 --opt overlap_trap_blocks=2" "--opt deep_thr=16 $D --opt deep_overlap=1 --opt overlap_trap_blocks=2" "--opt deep_thr=30 $D --opt deep_overlap=1"
mv gpurun_out/ab.log gpurun_out/${T}_ab.log
