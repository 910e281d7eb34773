# Session 5: compute-sanitizer (memcheck, initcheck) over the small GPU parity tests with the final kernels (new site-record layout,
# one-quad event record, call-free walk arithmetic, 7 / 8 blocks per SM)
mkdir -p gpurun_out; S=gpurun_out/r2ax_sanitizer_summary.log; rm -f $S
SEL='replay_of_reference_draws_is_bit_exact or philox_matches_oracle or golden_film_contacts or track_particle_replays or chunking_scheduling or shortcuts_do_not_change or contact_loop_replays or trace_capacity or one_warp_per_row or midpoint or host_state_refuses'
for tool in memcheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contacts.py -k "$SEL" -x -q > gpurun_out/r2ax_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> $S
  grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2ax_sanitizer_$tool.log | tail -3 >> $S
done
cat $S
