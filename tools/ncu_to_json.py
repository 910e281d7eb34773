"""profiles/round<R>_<workload>_hop_ncu.json from one `ncu --set full` capture of the hop kernel: the counters bench.py's
roofline block quotes (DRAM bytes per launch, L2 sectors), stamped with the hash of the kernel sources they were
captured from.  bench.py ignores a file whose hash differs from the sources it runs.

    python tools/ncu_to_json.py <file.ncu-rep> <workload> <excitons> <steps_per_launch> <hops_per_launch> [round]
"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import source_hash

rep, workload, excitons, steps, hops = sys.argv[1], sys.argv[2], int(float(sys.argv[3])), int(sys.argv[4]), float(sys.argv[5])
rnd = sys.argv[6] if len(sys.argv) > 6 else "2"
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
r = data[0]


def val(name, scale_units=True):
    if name not in hdr:
        return None
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    if scale_units:
        for pre, f in (("gbyte", 1e9), ("mbyte", 1e6), ("kbyte", 1e3), ("msecond", 1e-3), ("usecond", 1e-6), ("nsecond", 1e-9), ("ms", 1e-3), ("us", 1e-6), ("ns", 1e-9)):
            if u == pre:
                return v * f
    return v


t = val("gpu__time_duration.sum")
dram = (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0)
sectors = val("lts__t_sectors.sum", False)
if sectors is None:
    sectors = sum(val(n, False) or 0 for n in ("lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum"))
doc = {
    "kernel": r[hdr.index("Kernel Name")][:80], "workload": workload, "excitons": excitons, "steps_per_launch": steps,
    "hops_per_launch": hops, "source_hash": source_hash(), "time_s_under_ncu": t,
    "dram_bytes_per_launch": dram, "dram_gbs": dram / t / 1e9,
    "l2_sectors": sectors, "l2_gbs": sectors * 32 / t / 1e9, "l2_sectors_per_hop": sectors / hops, "dram_bytes_per_hop": dram / hops,
    "l2_hit_pct": val("lts__t_sector_hit_rate.pct", False), "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct", False),
    "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio", False),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
    "registers": val("launch__registers_per_thread", False), "warp_instructions": val("smsp__inst_executed.sum", False),
    "source": os.path.basename(rep),
}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "round%s_%s_hop_ncu.json" % (rnd, workload))
with open(path, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps(doc))
