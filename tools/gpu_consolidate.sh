set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r44_smoke.log 2>&1; echo smoke rc=$? >> gpurun_out/r44_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r44_pytest_gpu.log 2>&1
timeout 600 python bench.py > gpurun_out/r44_bench_default.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/r44_bench_reference.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r44_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r44_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kubo_kernel -s 1 -c 1 -o gpurun_out/r44_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r44_ncu_bench.log 2>&1
tail -3 gpurun_out/r44_smoke.log; tail -3 gpurun_out/r44_pytest_gpu.log; tail -1 gpurun_out/r44_bench_default.log | cut -c1-300; tail -1 gpurun_out/r44_bench_reference.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline --excitons 4000000 2>&1 | tail -1 > gpurun_out/r44_bench_4e6.log
timeout 300 python tools/run_c5.py 2000000 100 25 2>&1 | tail -1 > gpurun_out/r44_c5.log
timeout 600 python tools/run_c4.py 1.0 1000000 2>&1 | tail -1 > gpurun_out/r44_c4.log
cut -c1-400 gpurun_out/r44_bench_4e6.log gpurun_out/r44_c5.log gpurun_out/r44_c4.log
