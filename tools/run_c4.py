"""BASELINE config 4 at full size: 20 000 tubes x 250 sites (5e6 sites, ~1.4e9 table entries, ~22 GB of CSR rows): the
HBM-gather regime.  Prints table-build time, table size and hop throughput.  Needs a B200 (about 40 GB of HBM)."""
import json, sys, time, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0   # fraction of the tubes (box shrinks with sqrt(scale))
P = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
cfg = dict(film.CONFIG_FILMS["C4"])
cfg["NT"] = int(cfg["NT"] * scale)
cfg["LX"] = cfg["LX"] * scale ** 0.5
t0 = time.time(); pos, ori = film.film(**cfg); t_film = time.time() - t0
e = Engine(mc_block(P)); e.set_mesh(pos, ori)
t0 = time.time(); e.kubo_init(); t_init = time.time() - t0
nnz = int(e.csr_nnz()) if hasattr(e, "csr_nnz") else None
e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 64, want_msd=False)            # warm-up
h0 = e.hops(); e.kubo_step(DT, 64, want_msd=False)
ms = e.last_step_ms(); hops = e.hops() - h0
print(json.dumps({"config": "C4 x %.3g" % scale, "sites": e.num_sites(), "excitons": P, "film_s": round(t_film, 2), "init_s": round(t_init, 2),
                  "csr_build_s": round(e.csr_build_seconds(), 3), "hops_per_s": hops / (ms * 1e-3), "ms_per_64_steps": ms,
                  "hops_per_exciton_step": hops / (64 * P), "midpoint_guards": e.csr_midpoint_guards()}))
