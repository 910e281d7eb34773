# A/B of engine options on the bench workload: bash tools/gpu_ab.sh "<bench args A>" "<bench args B>" ...
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
run() { echo "== $*" >> gpurun_out/ab.log; timeout 200 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline $* 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> gpurun_out/ab.log 2>&1; }
for a in "$@"; do run $a; done
cat gpurun_out/ab.log
