# compute-sanitizer passes over the small GPU parity tests (memcheck, racecheck, initcheck); logs under gpurun_out/
mkdir -p gpurun_out
SEL='replay_of_reference_draws_is_bit_exact or philox_matches_oracle or golden_film_contacts or track_particle_replays or chunking_sorting or shortcuts_do_not_change'
for tool in memcheck racecheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contacts.py -k "$SEL" -x -q > gpurun_out/r42_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/r42_sanitizer_summary.log
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r42_sanitizer_$tool.log | tail -3 >> gpurun_out/r42_sanitizer_summary.log
done
cat gpurun_out/r42_sanitizer_summary.log
