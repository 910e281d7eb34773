"""BASELINE config 5 (per-GPU share): contact-driven transport on the C2 film -- monte_carlo::init, then
step / save_metrics / repopulate_contacts iterations.  Prints exciton population, hop throughput and the steady-state
current profile.      python tools/run_c5.py [c1_pop] [iterations] [steps_per_call]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT

c1 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 100
per_call = int(sys.argv[3]) if len(sys.argv) > 3 else 25
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(1)); e.set_mesh(pos, ori)
t0 = time.time(); e.init(c1, 0, seed=1, capacity=int(7 * c1)); t_init = time.time() - t0
P0 = e.number_of_particles()
e.step(DT, per_call)                                   # warm-up
h0 = e.hops(); ms = 0.0; t0 = time.time(); pops = []; curs = []
for _ in range(iters // per_call):
    pop, cur = e.step(DT, per_call); ms += e.last_step_ms(); pops.append(pop); curs.append(cur)
wall = time.time() - t0
hops = e.hops() - h0
pop = np.concatenate(pops); cur = np.concatenate(curs)
print(json.dumps({"config": "C5 share: C2 film, contacts, c1_pop %d" % c1, "initial_excitons": P0, "excitons_now": e.number_of_particles(),
                  "init_s": round(t_init, 2), "iterations": int(len(pop)), "device_ms_per_iteration": ms / len(pop), "wall_ms_per_iteration": wall * 1e3 / len(pop),
                  "hops_per_s_device": hops / (ms * 1e-3), "hops_per_s_wall": hops / wall, "excitons_counted_per_iteration": float(pop.sum(1).mean()),
                  "mean_net_crossings": [float(x) for x in cur.mean(0)]}))
