# compute-sanitizer passes over the small GPU parity tests (memcheck, racecheck, initcheck); logs under gpurun_out/
mkdir -p gpurun_out; rm -f gpurun_out/r2_sanitizer_summary.log
SEL='replay_of_reference_draws_is_bit_exact or philox_matches_oracle or golden_film_contacts or track_particle_replays or chunking_scheduling or shortcuts_do_not_change or contact_loop_replays or trace_capacity or one_warp_per_row or midpoint or host_state_refuses'
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contacts.py -k "$SEL" -x -q > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r2_sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_sanitizer_$tool.log | tail -3 >> gpurun_out/r2_sanitizer_summary.log
done
cat gpurun_out/r2_sanitizer_summary.log
