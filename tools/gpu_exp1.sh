mkdir -p gpurun_out
for v in FAKE_LOG FAKE_DIV BOTH; do
  echo "== $v" >> gpurun_out/r21_exp.log
  CNTMC_LIB=$PWD/exp/libcntmc_$v.so timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])" >> gpurun_out/r21_exp.log 2>&1
done
echo "== baseline" >> gpurun_out/r21_exp.log
timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'])" >> gpurun_out/r21_exp.log 2>&1
for P in 4000000 12500000; do for smb in 4096 32768; do
  echo "== P $P stage_mb $smb" >> gpurun_out/r21_exp.log
  timeout 300 python bench.py --excitons $P --stage-mb $smb --steps 3 --warmup 3 --e2e-steps 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['gpu_launches'])" >> gpurun_out/r21_exp.log 2>&1
done; done
cat gpurun_out/r21_exp.log
