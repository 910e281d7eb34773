"""Join ncu's per-SASS-instruction counters with nvdisasm's line table and report the hottest source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [N]
"""
import csv, glob, os, re, subprocess, sys, tempfile
rep, key = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(ROOT, 'cnt_film_monte_carlo_b200', 'libcntmc.so')], cwd=tmp, capture_output=True)
lines = []  # (file:line, inline-stack, sass text) in program order for the chosen function
for cubin in glob.glob(os.path.join(tmp, '*.cubin')):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    cur, infn = None, False
    for ln in txt.splitlines():
        m = re.match(r'\s*\.text\.(\S+):', ln)
        if m:
            infn = key in m.group(1); continue
        if not infn: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', ln)
        if m: lines.append((cur, m.group(1)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = next(i for i, r in enumerate(rows[:10]) if 'Source' in r)
hdr, data = rows[h], rows[h + 1:]
ix = {n: hdr.index(n) for n in ['Source', '# Samples', 'Instructions Executed', 'Thread Instructions Executed']}
assert len(data) == len(lines), (len(data), len(lines))
agg = {}
tot_i = tot_s = 0
for r, (loc, sass) in zip(data, lines):
    ie, te, sm = int(r[ix['Instructions Executed']] or 0), int(r[ix['Thread Instructions Executed']] or 0), int(r[ix['# Samples']] or 0)
    k = (loc[0], loc[1]) if loc else ('?', 0)
    a = agg.setdefault(k, [0, 0, 0]); a[0] += ie; a[1] += te; a[2] += sm
    tot_i += ie; tot_s += sm
print('total warp-instructions %.3e  samples %d' % (tot_i, tot_s))
src = {}
for (f, l), (ie, te, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    if f not in src:
        p = os.path.join(ROOT, 'cnt_film_monte_carlo_b200', 'csrc', f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][l - 1].strip() if 0 < l <= len(src[f]) else ''
    print('%5.1f%% inst %5.1f%% samp thr/inst %4.1f  %s:%d  %s' % (100 * ie / tot_i, 100 * sm / max(tot_s, 1), te / max(ie, 1), f, l, text[:90]))
