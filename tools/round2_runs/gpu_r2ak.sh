# Session 5: are the spills of the 6-blocks-per-SM instantiation what it loses by, or the sixth block itself?
# The 80-register kernel run with only 5 (and 4) blocks per SM resident, against the 96-register kernel at 5; and the same with a
# call-free square root in the chain walk (libcntmc_fsq.so: 88 instead of 104 bytes of spill stores at 80 registers).
mkdir -p gpurun_out; L=gpurun_out/r2ak_ab.log; rm -f $L
run() { echo "== $*" >> $L; env $1 $2 timeout 300 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline ${@:3} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_launch'], d['e2e']['value'], d['roofline']['frac'])" >> $L 2>&1; }
F=CNTMC_LIB=$PWD/cnt_film_monte_carlo_b200/libcntmc_fsq.so
N=CNTMC_X=0
run $N CNTMC_Y=0
run $N CNTMC_Y=0 --occupancy 6
run $N CNTMC_DBG_BLOCKS_PER_SM=5 --occupancy 6
run $N CNTMC_DBG_BLOCKS_PER_SM=4 --occupancy 6
run $N CNTMC_DBG_BLOCKS_PER_SM=4
run $F CNTMC_Y=0
run $F CNTMC_Y=0 --occupancy 6
run $F CNTMC_DBG_BLOCKS_PER_SM=5 --occupancy 6
run $N CNTMC_Y=0 --workload C4
run $N CNTMC_DBG_BLOCKS_PER_SM=5 --workload C4 --occupancy 6
run $F CNTMC_Y=0 --workload C4
run $F CNTMC_Y=0 --workload C4 --occupancy 6
cat $L
