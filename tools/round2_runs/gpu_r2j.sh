# ncu evidence, round 2: hop kernel on C2 and C4, table build on C4, launch list of the default bench
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/r2_c2_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_c2_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/r2_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_c4_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:csr_rows_kernel -c 2 -o gpurun_out/r2_c4_csr -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_c4_csr_bench.log 2>&1
timeout 900 $NCU -k regex:csr_rows_kernel -c 2 -o gpurun_out/r2_c2_csr -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_c2_csr_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_launch_bench.log 2>&1
ls -la gpurun_out/r2_*.ncu-rep
tail -1 gpurun_out/r2_c2_ncu_bench.log | cut -c1-300
tail -1 gpurun_out/r2_c4_ncu_bench.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k midpoint 2>&1 | tail -3
