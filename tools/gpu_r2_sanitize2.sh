#!/bin/bash
# compute-sanitizer over what the last session of round 2 added: the davoody placement kernel and the trap kernel beside the lane
# kernel (in-launch hand-over); memcheck and initcheck (racecheck only sees shared memory inside one kernel)
mkdir -p gpurun_out; rm -f gpurun_out/r2b_sanitizer_summary.log
for tool in memcheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_davoody.py tests/test_gpu_parity.py -k "first_order_matches_the_reference_vectors or engine_builds_its_davoody or chunking_scheduling" -x -q > gpurun_out/r2b_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2b_sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2b_sanitizer_$tool.log | tail -3 >> gpurun_out/r2b_sanitizer_summary.log
done
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_davoody.py -k "first_order_matches_the_reference_vectors" -x -q > gpurun_out/r2b_sanitizer_racecheck.log 2>&1
echo "racecheck (davoody) rc=$?" >> gpurun_out/r2b_sanitizer_summary.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/r2b_sanitizer_racecheck.log | tail -3 >> gpurun_out/r2b_sanitizer_summary.log
cat gpurun_out/r2b_sanitizer_summary.log
# the in-launch hand-over alone, three times (an initcheck report that comes and goes is the tool's shadow memory racing between
# the two concurrent kernels, not the data: the results are compared bit for bit in the same run)
for i in 1 2 3; do
  timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python tools/overlap_check.py > gpurun_out/r2b_sanitizer_overlap_init$i.log 2>&1
  echo "overlap initcheck run $i rc=$? $(grep -E 'ERROR SUMMARY|overlap ok' gpurun_out/r2b_sanitizer_overlap_init$i.log | tr '\n' ' ')" >> gpurun_out/r2b_sanitizer_summary.log
done
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/overlap_check.py > gpurun_out/r2b_sanitizer_overlap_mem.log 2>&1
echo "overlap memcheck rc=$? $(grep -E 'ERROR SUMMARY|overlap ok' gpurun_out/r2b_sanitizer_overlap_mem.log | tr '\n' ' ')" >> gpurun_out/r2b_sanitizer_summary.log
tail -5 gpurun_out/r2b_sanitizer_summary.log
