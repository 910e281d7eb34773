python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
sed -n '/^# the in-launch/,$p' tools/gpu_r2_sanitize2.sh > /tmp/ov.sh; rm -f gpurun_out/r2b_sanitizer_summary.log; mkdir -p gpurun_out; bash /tmp/ov.sh
