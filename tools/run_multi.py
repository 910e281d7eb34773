"""The product's own multi-GPU path (cntmc_multi_*: one host thread, one handle per GPU, run-time NCCL) on the C2 workload:
N GPUs x 1e6 excitons, checks a shard against a single-GPU run of the same global ids, prints hops/s by wall clock.

    python tools/run_multi.py [n_gpus] [excitons_per_gpu] [calls]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine, MultiEngine
from bench import mc_block, DT

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 10
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
m = MultiEngine(mc_block(per * n), list(range(n)))
m.set_mesh(pos, ori)
t0 = time.time(); m.kubo_init(); t_init = time.time() - t0
m.kubo_create_particles(per * n, seed=1)
for _ in range(3):
    m.kubo_step(DT, 100)
h0 = m.hops(); t0 = time.perf_counter()
for _ in range(calls):
    msd = m.kubo_step(DT, 100)
sec = time.perf_counter() - t0
hops = m.hops() - h0
# the last shard against one GPU running the same global ids alone
first = per * (n - 1)
e = Engine(mc_block(per), device=0); e.set_mesh(pos, ori); e.kubo_init(); e.kubo_create_particles(per, seed=1, first_global_id=first)
for _ in range(3 + calls):
    e.kubo_step(DT, 100, want_msd=False)
pm, pe = m.particles(), e.particles()
same = all(np.array_equal(pm[k][..., first:first + per], pe[k]) for k in pe)
print(json.dumps({"path": "cntmc_multi (libcntmc.so, run-time NCCL %d)" % m.nccl_version(), "n_gpus": n, "excitons_total": per * n,
                  "init_s": round(t_init, 2), "hops_per_s_wall_clock": hops / sec, "ms_per_call": 1e3 * sec / calls,
                  "last_shard_equals_single_gpu_run": bool(same), "msd_last": msd[-1].tolist()}))
