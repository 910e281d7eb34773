mkdir -p gpurun_out; rm -f gpurun_out/r2g_*.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2g_pytest.log; cat gpurun_out/r2g_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline > gpurun_out/r2g_c2.log 2>&1; tail -1 gpurun_out/r2g_c2.log | cut -c1-1500
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline --opt host_slices=1 > gpurun_out/r2g_c2_noslice.log 2>&1; tail -1 gpurun_out/r2g_c2_noslice.log | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 5 --no-cpu-baseline --opt host_slices=8 > gpurun_out/r2g_c2_slice8.log 2>&1; tail -1 gpurun_out/r2g_c2_slice8.log | cut -c1-400
timeout 300 python bench.py --workload C1 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2g_c1.log 2>&1; tail -1 gpurun_out/r2g_c1.log | cut -c1-700
timeout 300 python bench.py --workload C3 --excitons-total 8000000 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2g_c3.log 2>&1; tail -1 gpurun_out/r2g_c3.log | cut -c1-700
timeout 600 python bench.py --workload C4 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2g_c4.log 2>&1; tail -1 gpurun_out/r2g_c4.log | cut -c1-900
timeout 600 python bench.py --workload C5 --c1-pop 2000000 --steps 3 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2g_c5.log 2>&1; tail -1 gpurun_out/r2g_c5.log | cut -c1-900
