"""Cycles per segment of the hop loop, lane 0 of every warp, one instrumented 64-step launch of a diagnostics build
(python -m cnt_film_monte_carlo_b200.build --segments; run with CNTMC_LIB=cnt_film_monte_carlo_b200/libcntmc_seg.so)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(P)); e.set_mesh(pos, ori)
for k in ("hot_pct", "chunk_steps", "occupancy"):
    if os.environ.get(k.upper()):
        e.set_option(k, int(os.environ[k.upper()]))
e.kubo_init(); e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 320, want_msd=False)
path = "/tmp/warp_times.bin"
os.environ["CNTMC_DEBUG_WARP_TIMES"] = path
e.set_option("stats", 1)
e.kubo_step(DT, 64, want_msd=False)
raw = np.fromfile(path, dtype=np.uint64).reshape(-1, 12)
role = (raw[:, 3] & np.uint64(1)).astype(int)
iters = (raw[:, 3] >> np.uint64(8)).astype(np.float64)
seg = raw[:, 4:].astype(np.float64)
names = ["E: fly", "E: hop info + draw + dice", "E: select (row probes)", "E: set_site / move", "E: ff (draw, log, 1/Gamma of dest)",
         "S: step-end path", "loop head / refill", "wait for the warp after own event"]
out = {"P": P, "call_ms": e.last_step_ms(), "iterations_mean": float(iters.mean())}
for r, nm in ((1, "hot"), (0, "cold")):
    m = role == r
    tot = seg[m].sum()
    out[nm] = {"warps": int(m.sum()), "cycles_per_iteration": float(seg[m].sum() / iters[m].sum()),
               "share": {n: round(float(seg[m][:, k].sum() / tot), 4) for k, n in enumerate(names)},
               "cycles_per_iteration_by_segment": {n: round(float(seg[m][:, k].sum() / iters[m].sum()), 1) for k, n in enumerate(names)}}
print(json.dumps(out))
