#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_davoody.py -x -q 2>&1 | tail -15 > gpurun_out/f1a_pytest.log
cat gpurun_out/f1a_pytest.log
timeout 600 python tools/davoody_bench.py --ref 2>&1 | tee gpurun_out/f1a_bench.log
