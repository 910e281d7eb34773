"""Per-source-line share of executed instructions / stall samples from an ncu report (needs -lineinfo + --import-source)."""
import csv, subprocess, sys
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--print-source', 'cuda', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr, data = rows[h], rows[h + 1:]
ci = {n: hdr.index(n) for n in ['Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed']}
num = lambda r, k: int(r[ci[k]] or 0) if len(r) > ci[k] else 0
tot_i = sum(num(r, 'Instructions Executed') for r in data)
tot_s = sum(num(r, '# Samples') for r in data)
print('total warp-instructions %.3e, samples %d' % (tot_i, tot_s))
top = sorted(data, key=lambda r: -num(r, 'Instructions Executed'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]
for r in top:
    ie, te = num(r, 'Instructions Executed'), num(r, 'Thread Instructions Executed')
    print('%5.1f%% inst %5.1f%% samp thr/inst %4.1f | %s' % (100 * ie / tot_i, 100 * num(r, '# Samples') / max(tot_s, 1), te / max(ie, 1), r[ci['Source']].strip()[:120]))
