# Session 5: the 7-blocks-per-SM default -- hot share and host slices re-tuned, one full ncu capture of the hop kernel on C2
mkdir -p gpurun_out; L=gpurun_out/r2am_ab.log; rm -f $L
run() { echo "== $*" >> $L; timeout 400 python bench.py --steps 10 --warmup 3 --e2e-steps 2 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
run
run --hot-pct 20
run --hot-pct 40
run --hot-pct 50
run --opt host_slices=2 --opt slice_share=2
run --opt host_slices=3 --opt slice_share=3
run --opt host_slices=7 --opt slice_share=7
run --chunk 100
cat $L
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:kubo_kernel -s 3 -c 1 -o gpurun_out/r2am_c2_kubo -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2am_c2_ncu_bench.log 2>&1
ls -la gpurun_out/r2am_c2_kubo.ncu-rep
