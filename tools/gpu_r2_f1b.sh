#!/bin/bash
# round 2, f1: the whole GPU suite (davoody tests included) and the table-build bench
mkdir -p gpurun_out; T=f1b
timeout 1500 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/${T}_pytest_gpu.log 2>&1
tail -12 gpurun_out/${T}_pytest_gpu.log
timeout 600 python tools/davoody_bench.py 2>&1 | tee gpurun_out/${T}_bench.log | cut -c1-400
