"""Generate tests/golden/small_forster_contacts_replay.npz: the reference's contact loop (monte_carlo::init, then step /
save_metrics / repopulate_contacts, main.cpp:98-106) on the small_forster film, run by the reference's own code with
every rand() draw attributed to the exciton that consumed it (oracle/t0_driver.cpp: t0_contact_iteration_logged, proven
equal to the verbatim loop by tests/test_oracle_t0.py).  Exciton ids are in order of birth.

    python tests/golden/make_golden_contacts.py        # dev container only
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cnt_film_monte_carlo_b200 import film  # noqa: E402
from oracle import t0 as T0m  # noqa: E402
from oracle import t1 as T1m  # noqa: E402
from conftest import Golden  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ITERATIONS, SEED = 12, 100


def main():
    g = Golden("small_forster")
    with tempfile.TemporaryDirectory() as tmp:
        mesh = os.path.join(tmp, "mesh")
        film.write_mesh(mesh, g.pos_nm, g.orient)
        mc = dict(g.mc)
        mc.update({"mesh input directory": mesh, "output directory": os.path.join(tmp, "out"), "keep old results": False})
        jpath = os.path.join(tmp, "input.json")
        with open(jpath, "w") as f:
            json.dump({"exciton monte carlo": mc}, f)
        t = T0m.T0()
        p0 = t.open_contacts_logged(jpath, SEED)
        for _ in range(ITERATIONS):
            t.contact_iteration_logged(g.dt)
        n_ids = t.next_id()
        off, flat = t.draws(n_ids)
        p = t.particles()
        out = dict(iterations=ITERATIONS, seed=SEED, dt=g.dt, c1_pop=1100, c2_pop=0, p0=p0, draw_off=off, draws=flat,
                   draw_logs=T1m.log_ratios(flat), ids=t.particle_ids(), site=p["site"], pos=p["pos"], ff=p["ff"], heading=p["heading"])
        t.log_draws(False)
        t.close()
        for name, key in (("population_profile.dat", "pop_file"), ("region_current.dat", "curr_file")):
            with open(os.path.join(tmp, "out", name)) as f:
                out[key] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "small_forster_contacts_replay.npz"), **out)
    print("ids", n_ids, "alive", len(out["ids"]), "draws", len(flat), "bytes", os.path.getsize(os.path.join(HERE, "small_forster_contacts_replay.npz")))


if __name__ == "__main__":
    main()
