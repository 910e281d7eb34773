mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r23_pytest_gpu.log 2>&1; tail -5 gpurun_out/r23_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:csr_rows -c 4 --csv --log-file gpurun_out/r23_csr_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > /dev/null 2>&1
grep csr_rows gpurun_out/r23_csr_launches.csv | cut -c1-300
timeout 600 python tools/run_c4.py 1.0 1000000 > gpurun_out/r23_c4_full.log 2>&1; tail -1 gpurun_out/r23_c4_full.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_contacts.py -k "replay_of_reference_draws_is_bit_exact or philox_matches_oracle or golden_film_contacts or track_particle_replays" -x -q > gpurun_out/r23_sanitizer_initcheck.log 2>&1; echo initcheck rc=$?; grep "ERROR SUMMARY\|passed" gpurun_out/r23_sanitizer_initcheck.log | tail -2
