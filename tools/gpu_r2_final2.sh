# final refresh for the committed sources: smoke, GPU tests, the two bench arms, launch list and the two full captures
set -x
mkdir -p gpurun_out; T=r2z
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/${T}_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_default.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.log 2>&1
timeout 900 python bench.py --workload C4 > gpurun_out/${T}_bench_C4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:kubo_kernel -s 3 -c 1 -o gpurun_out/${T}_c2_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c2_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/${T}_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:csr_fill_warp_kernel -c 1 -o gpurun_out/${T}_c4_csr_warp -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_csr_bench.log 2>&1
set +x
tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest_gpu.log
for f in default reference C4; do grep -E '^\{"(metric|impl)' gpurun_out/${T}_bench_$f.log | tail -1 | cut -c1-260; done
