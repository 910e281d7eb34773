"""Distribution of exciton activity on the bench workload: expected events per step Gamma(site)*dt of every exciton after
a warm-up, and measured events per exciton over one 64-step launch."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(P)); e.set_mesh(pos, ori); e.kubo_init(); e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 300, want_msd=False)
gam = e.sites()["max_rate"]
p = e.particles()
act = gam[p["site"]] * DT
e.trace_enable(1)                      # per-exciton event counts of the next call
e.kubo_step(DT, 64, want_msd=False)
counts, _ = e.trace()
out = {"P": P}
for name, v in (("expected_events_per_step", act), ("events_per_64_steps", counts.astype(np.float64))):
    out[name] = {"mean": float(v.mean()), "pct": {str(q): float(np.percentile(v, q)) for q in (50, 90, 99, 99.9, 99.99, 100)}}
for thr in (8, 16, 32, 64, 128):
    m = act >= thr
    out["act>=%d" % thr] = {"excitons": int(m.sum()), "share_of_events": float(counts[m].sum() / counts.sum())}
for thr in (512, 1024, 2048, 3072, 4096, 6144):
    m = counts >= thr
    out["events>=%d" % thr] = {"excitons": int(m.sum()), "share_of_events": float(counts[m].sum() / counts.sum())}
# how well does the state at launch start predict the launch's events?
hot = counts >= 2048
out["hot_predicted_by_act>=16"] = float((act[hot] >= 16).mean()) if hot.any() else None
out["distinct_sites_of_hot"] = int(len(np.unique(p["site"][hot])))
print(json.dumps(out))
