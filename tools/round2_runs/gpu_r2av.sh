# Session 5: full ncu captures of the plain hop kernel for the final sources (C2 at 7 blocks per SM, C4 at 8), picked by the demangled
# template arguments; tools/ncu_to_json.py turns them into the hash-stamped counters bench.py quotes
mkdir -p gpurun_out; T=r2av
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 900 $NCU -k 'regex:kubo_kernel<.*bool.0, .bool.0>' -s 2 -c 1 -o gpurun_out/${T}_c2_kubo -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c2_ncu_bench.log 2>&1
timeout 900 $NCU -k 'regex:kubo_kernel<.*bool.0, .bool.0>' -s 1 -c 1 -o gpurun_out/${T}_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_ncu_bench.log 2>&1
ls -la gpurun_out/${T}_*.ncu-rep
