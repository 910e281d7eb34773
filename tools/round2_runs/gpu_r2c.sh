mkdir -p gpurun_out; rm -f gpurun_out/r2c_diag.log
for a in "1000000 16" "1000000 8" "1000000 30" "4000000 16"; do timeout 200 python tools/deep_diag.py $a >> gpurun_out/r2c_diag.log 2>&1; done
cat gpurun_out/r2c_diag.log
