mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_c5_launches.csv python bench.py --workload C5 --c1-pop 4000000 --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2p_c5_bench.log 2>&1
python - <<'PY'
import csv, collections
rows=list(csv.reader(open('gpurun_out/r2p_c5_launches.csv')))
hdr=None; tot=collections.Counter(); cnt=collections.Counter()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        name=r[hdr.index('Kernel Name')].split('(')[0][:70]; v=float(r[hdr.index('Metric Value')].replace(',','')); u=r[hdr.index('Metric Unit')]
        v = v*{'ns':1e-6,'us':1e-3,'ms':1.0,'nsecond':1e-6,'usecond':1e-3,'msecond':1.0}.get(u,1e-6)
        tot[name]+=v; cnt[name]+=1
T=sum(tot.values())
for k,v in tot.most_common(12): print('%-72s %4d launches %9.3f ms %5.1f%%' % (k,cnt[k],v,100*v/T))
PY
