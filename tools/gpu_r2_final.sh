# round 2 evidence run (one B200): smoke, GPU tests, the driver's two bench arms, every workload, ncu launch list + full captures
set -x
mkdir -p gpurun_out; T=r2f
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/${T}_smoke.log
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/${T}_pytest_gpu.log 2>&1
timeout 900 python bench.py > gpurun_out/${T}_bench_default.log 2>&1
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_reference.log 2>&1
for w in C1 C4 C5; do timeout 1200 python bench.py --workload $w --steps 5 --warmup 2 --e2e-steps 2 > gpurun_out/${T}_bench_$w.log 2>&1; done
timeout 1800 python bench.py --workload C3 --steps 3 --warmup 1 --e2e-steps 1 > gpurun_out/${T}_bench_C3.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline --excitons 4000000 > gpurun_out/${T}_bench_4e6.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_launch_bench.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/${T}_c2_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c2_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/${T}_c4_kubo -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_ncu_bench.log 2>&1
timeout 900 $NCU -k regex:csr_rows_kernel -c 2 -o gpurun_out/${T}_c4_csr -f python bench.py --workload C4 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c4_csr_bench.log 2>&1
timeout 900 $NCU -k regex:contact_kernel -s 1 -c 1 -o gpurun_out/${T}_c5_contact -f python bench.py --workload C5 --c1-pop 2000000 --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/${T}_c5_ncu_bench.log 2>&1
set +x
tail -2 gpurun_out/${T}_smoke.log; tail -3 gpurun_out/${T}_pytest_gpu.log
for f in default reference C1 C3 C4 C5 4e6; do tail -1 gpurun_out/${T}_bench_$f.log | cut -c1-260; done
