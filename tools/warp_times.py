"""Per-warp timeline of one instrumented 64-step launch (CNTMC_DEBUG_WARP_TIMES): when warps first fail to get work and when they exit."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
hot = int(sys.argv[2]) if len(sys.argv) > 2 else 40
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(P)); e.set_mesh(pos, ori); e.set_option("hot_pct", hot)
e.kubo_init(); e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 320, want_msd=False)
e.kubo_step(DT, 64, want_msd=False)
path = "/tmp/warp_times.bin"
os.environ["CNTMC_DEBUG_WARP_TIMES"] = path
e.set_option("stats", 1)
e.kubo_step(DT, 64, want_msd=False)
raw = np.fromfile(path, dtype=np.uint64).reshape(-1, 4)
w = raw.astype(np.float64)
t0 = w[:, 0].min()
dry, ex, role = (w[:, 1] - t0) / 1e6, (w[:, 2] - t0) / 1e6, (raw[:, 3] & np.uint64(1)).astype(np.float64)
iters = (raw[:, 3] >> np.uint64(8)).astype(np.float64)
out = {"P": P, "hot_pct": hot, "span_ms": float(ex.max())}
for name, m in (("hot", role == 1), ("cold", role == 0)):
    out[name] = {"warps": int(m.sum()), "first_dry_ms_pct": [float(np.percentile(dry[m], q)) for q in (0, 10, 50, 90, 100)],
                 "exit_ms_pct": [float(np.percentile(ex[m], q)) for q in (0, 10, 50, 90, 99, 100)],
                 "iterations_pct": [float(np.percentile(iters[m], q)) for q in (10, 50, 90, 100)],
                 "us_per_iteration_pct": [float(np.percentile(1e3 * ex[m] / np.maximum(iters[m], 1), q)) for q in (10, 50, 90)]}
print(json.dumps(out))
