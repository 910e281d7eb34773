# Session 5: what the box's host link can do (tools/pcie_probe.py) against what the end-to-end step pays, and the slice settings again
mkdir -p gpurun_out; L=gpurun_out/r2aw_ab.log; rm -f $L
python tools/pcie_probe.py 2>&1 | tail -1 | tee gpurun_out/r2aw_pcie.json
run() { echo "== $*" >> $L; timeout 400 python bench.py --steps 5 --warmup 3 --e2e-steps 6 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value']/d['value'])" >> $L 2>&1; }
run --opt host_slices=1
run --opt host_slices=2 --opt slice_share=2
run --opt host_slices=4 --opt slice_share=4
run --opt host_slices=4 --opt slice_share=3
run --opt host_slices=6 --opt slice_share=6
run --opt host_slices=8 --opt slice_share=8
run --opt host_slices=16 --opt slice_share=16
cat $L
