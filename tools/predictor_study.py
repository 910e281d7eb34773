"""How well can the activity of an exciton in the next launch be predicted at its start?  (CPU study on the C2 film.)"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from emul import Emul
from cnt_film_monte_carlo_b200 import film
from conftest import base_mc

P = 20000
e = Emul(base_mc()); e.kubo_init(*film.film(**film.CONFIG_FILMS["C2"]))
gam = e.sites()["max_rate"]
e.create_philox(P, seed=1)
e.kubo_step(1e-13, 64)                      # warm up
def events_in(nsteps):
    h0 = e.particles()["ndraw"].astype(np.int64).copy()
    s0 = e.particles()["site"].copy()
    e.kubo_step(1e-13, nsteps)
    d = e.particles()["ndraw"].astype(np.int64) - h0
    return s0, d // 2                         # ~2 draws per event
for chunk in (8, 16, 64):
    s_prev, ev_prev = events_in(chunk)
    s0, ev = events_in(chunk)
    g = gam[s0] * 1e-13                       # expected events per step if it stayed on this site
    for name, key in (("events in previous chunk", ev_prev), ("Gamma(site)*dt at chunk start", g), ("max of both", np.maximum(ev_prev / chunk, g))):
        order = np.argsort(-key, kind="stable")
        top = order[: P // 10]
        # how much of the next chunk's events do the top-10% predicted excitons hold, vs the true top 10%
        true_top = np.argsort(-ev)[: P // 10]
        print("chunk %2d  %-32s rank-corr %.3f  events captured by predicted top 10%%: %.1f%% (oracle %.1f%%); hottest missed: %d" % (
            chunk, name, np.corrcoef(np.argsort(np.argsort(key)), np.argsort(np.argsort(ev)))[0, 1],
            100 * ev[top].sum() / max(ev.sum(), 1), 100 * ev[true_top].sum() / max(ev.sum(), 1), ev[order[P // 10:]].max()))
