"""The attributed loops of oracle/t0_driver.cpp against the reference's own loops, live (dev container only: needs
oracle/_ref/libt0.so = the reference's translation units compiled from /root/reference).  The golden fixtures were
generated through the attributed loops; this proves they are the verbatim loops, bit for bit."""
import json
import os

import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from oracle import t0 as T0m

pytestmark = pytest.mark.skipif(not T0m.available(), reason="oracle/_ref/libt0.so not built (no /root/reference here)")


def write_case(tmp_path, g, out):
    mesh = str(tmp_path / "mesh")
    film.write_mesh(mesh, g.pos_nm, g.orient)
    mc = dict(g.mc)
    mc.update({"mesh input directory": mesh, "output directory": str(tmp_path / out), "keep old results": False})
    path = str(tmp_path / (out + ".json"))
    with open(path, "w") as f:
        json.dump({"exciton monte carlo": mc}, f)
    return path, mc


def test_logged_kubo_loop_is_the_verbatim_loop(tmp_path, golden_small):
    g = golden_small
    path, _ = write_case(tmp_path, g, "out")
    t = T0m.T0()
    t.open(path, g.seed)
    t.srand(g.seed)
    t.create_particles_verbatim()
    msd_v = t.kubo_step_verbatim(g.dt, g.nsteps)
    pv = t.particles()
    t.open(path, g.seed)
    t.log_draws(True)
    t.srand(g.seed)
    t.create_particles_logged(g.P)
    msd_l = t.kubo_step_logged(g.dt, g.nsteps)
    pl = t.particles()
    t.log_draws(False)
    t.close()
    assert np.array_equal(msd_v, msd_l) and np.array_equal(msd_l, g.z["msd"])
    assert all(np.array_equal(pv[k], pl[k]) for k in pv)


def test_logged_contact_loop_is_the_verbatim_loop(tmp_path, golden_small):
    g = golden_small
    res = []
    for logged in (False, True):
        path, mc = write_case(tmp_path, g, "out_%d" % logged)
        t = T0m.T0()
        if logged:
            p0 = t.open_contacts_logged(path, g.seed)
            assert p0 == 5 * 1100
        else:
            t.log_draws(False)
            t.open_contacts(path, g.seed)
        for _ in range(15):
            (t.contact_iteration_logged if logged else t.contact_iteration)(g.dt)
        p = t.particles()
        files = {}
        for name in ("population_profile.dat", "region_current.dat"):
            with open(os.path.join(mc["output directory"], name)) as f:
                files[name] = f.read()
        if logged:
            ids = t.particle_ids()
            off, flat = t.draws(t.next_id())
            assert t.next_id() == 5500 + 15 * 1100 and len(ids) == len(p["site"]) and len(np.unique(ids)) == len(ids)
            assert off[-1] == len(flat) and np.all(np.diff(off) >= 3)     # site, free flight, heading at birth; two per event
        t.log_draws(False)
        t.close()
        res.append((p, files))
    assert all(np.array_equal(res[0][0][k], res[1][0][k]) for k in res[0][0])
    assert res[0][1] == res[1][1]
