#!/bin/bash
# validation of the committed state: smoke, GPU suite, default bench, table bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py 2>&1 | grep '^{"metric' | tail -1 > gpurun_out/r2ac_bench_default.json; python -c "
import json; d=json.load(open('gpurun_out/r2ac_bench_default.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['roofline'].get('l2'), d['clocks'])"
timeout 600 python tools/davoody_bench.py 2>&1 | cut -c1-220
