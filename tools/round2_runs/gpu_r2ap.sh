# Session 5: full ncu capture of the plain (not instrumented) hop kernel on C2 -- with one 100-step launch per bench step the
# fourth kubo_kernel launch is the instrumented counting pass, so the kernel is picked by its demangled template arguments
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
timeout 900 $NCU -k 'regex:kubo_kernel<.*bool.0, .bool.0>' -s 2 -c 1 -o gpurun_out/r2ap_c2_kubo -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2ap_c2_ncu_bench.log 2>&1
ls -la gpurun_out/r2ap_c2_kubo.ncu-rep; grep -c metric gpurun_out/r2ap_c2_ncu_bench.log
