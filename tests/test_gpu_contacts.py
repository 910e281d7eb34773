"""Contact mode on the GPU (through the C ABI) against the oracle: populations and interface currents are integers and
must match exactly; so must the surviving population."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import CntmcError, Engine
from oracle import t1 as T1m
from conftest import base_mc

pytestmark = pytest.mark.gpu


def run_both(mc, pos, ori, c1, c2, seed, dt, nsteps, opts=()):
    e = Engine(mc)
    e.set_mesh(pos, ori)
    for k, v in opts:
        e.set_option(k, v)
    e.init(c1, c2, seed=seed)
    t = T1m.T1()
    t.draws_philox(seed)
    t.set_memo(True)
    t.contacts_init(mc, pos, ori, c1_pop=c1, c2_pop=c2)
    assert np.array_equal(e.area(), t.area())
    assert np.array_equal(e.contact_sites(1), t.contact_sites(1)) and np.array_equal(e.contact_sites(2), t.contact_sites(2))
    p0e, p0t = e.particles(), t.particles()
    assert np.array_equal(p0e["site"], p0t["site"]) and np.array_equal(p0e["pos"], p0t["pos"])
    hist = [t.contact_iteration(dt) for _ in range(nsteps)]
    pop_t, cur_t = np.array([h[0] for h in hist]), np.array([h[1] for h in hist])
    return e, t, pop_t, cur_t


def test_golden_film_contacts_match_oracle(golden_small):
    g = golden_small
    e, t, pop_t, cur_t = run_both(g.mc, g.pos_nm, g.orient, 1100, 0, 5, 1e-14, 60)
    assert np.array_equal(e.area(), g.z["contact_area"]) and np.array_equal(e.contact_sites(1), g.z["contact_c1"])
    pop, cur = e.step(1e-14, 60)
    assert np.array_equal(pop, pop_t) and np.array_equal(cur, cur_t)
    assert e.number_of_particles() == t.L.t1_num_particles(t.h)
    assert e.hops() == t.hops()
    # survivors: same set of excitons (order differs: the reference swaps to the tail, the engine compacts)
    pe, pt = e.particles(), t.particles()
    assert np.array_equal(np.sort(pe["site"]), np.sort(pt["site"]))
    assert e.time() == t.time()


def test_chunking_and_both_contacts(golden_small):
    g = golden_small
    e, t, pop_t, cur_t = run_both(g.mc, g.pos_nm, g.orient, 500, 120, 9, 2e-14, 45, opts=(("chunk_steps", 7),))
    got = [e.step(2e-14, n) for n in (1, 20, 24)]
    pop = np.concatenate([p for p, _ in got])
    cur = np.concatenate([c for _, c in got])
    assert np.array_equal(pop, pop_t) and np.array_equal(cur, cur_t)
    assert e.number_of_particles() == t.L.t1_num_particles(t.h)


def test_larger_film_contacts():
    pos, ori = film.film(NT=150, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
    mc = base_mc(**{"number of segments": 8})
    e, t, pop_t, cur_t = run_both(mc, pos, ori, 4000, 0, 3, 1e-13, 25)
    pop, cur = e.step(1e-13, 25)
    assert np.array_equal(pop, pop_t) and np.array_equal(cur, cur_t)
    assert e.number_of_particles() == t.L.t1_num_particles(t.h)


def format_path(path):
    """the rows monte_carlo::track_particle writes (monte_carlo.h:806-815): showpos, scientific, 6 digits"""
    return "".join("   %+.6e %+.6e %+.6e\n" % tuple(r) for r in path) + "\n"


def test_track_particle_replays_the_reference_trajectory(golden_small):
    """monte_carlo::track_particle: with the reference's own draws the path file is reproduced byte for byte."""
    g = golden_small
    e = Engine(g.mc)
    e.set_mesh(g.pos_nm, g.orient)
    e.init(0, 0)
    path, reached = e.track_particle(float(g.z["track_dt"]), replay_draws=g.z["track_draws"], replay_logs=g.z["track_logs"])
    assert reached
    assert format_path(path) == bytes(g.z["track_file"]).decode()
    t = T1m.T1()
    t.contacts_init(g.mc, g.pos_nm, g.orient)
    t.draws_replay(np.array([0, len(g.z["track_draws"])], np.int64), g.z["track_draws"])
    path_t, _ = t.track_particle(float(g.z["track_dt"]))
    assert np.array_equal(path, path_t)               # bit for bit, not only to the 7 printed digits


def test_track_particle_philox_matches_oracle(golden_small):
    g = golden_small
    e = Engine(g.mc)
    e.set_mesh(g.pos_nm, g.orient)
    e.init(0, 0)
    t = T1m.T1()
    t.draws_philox(21)
    t.set_memo(True)
    t.contacts_init(g.mc, g.pos_nm, g.orient, c1_pop=0, c2_pop=0)
    for gid in (0, 5):
        path, reached = e.track_particle(1e-12, seed=21, global_id=gid, max_steps=4000)
        path_t, reached_t = t.track_particle(1e-12, gid=gid, max_steps=4000)
        assert reached == reached_t and path.shape == path_t.shape and len(path) > 10
        assert np.allclose(path, path_t, rtol=1e-9, atol=1e-18)
    short, reached = e.track_particle(1e-12, seed=21, global_id=0, max_steps=7)     # the bound the reference lacks
    assert len(short) == 7 and not reached
    assert np.allclose(short, t.track_particle(1e-12, gid=0, max_steps=7)[0], rtol=1e-9, atol=1e-18)


def test_mirror_writes_statistics_and_path_files(golden_small, tmp_path):
    from cnt_film_monte_carlo_b200.monte_carlo import monte_carlo
    g = golden_small
    mesh, out = str(tmp_path / "mesh"), str(tmp_path / "out")
    film.write_mesh(mesh, g.pos_nm, g.orient)
    mc = dict(g.mc, **{"mesh input directory": mesh, "output directory": out})
    sim = monte_carlo(mc, quiet=True)
    sim.init()
    with open(out + "/scatterer_statistics.dat") as f:
        assert f.read() == bytes(g.z["contact_stat_file"]).decode()      # monte_carlo.h:691-719, the reference's own file
    assert sim.track_particle(2e-12, 3, max_steps=100000)
    with open(out + "/particle_path.3.dat") as f:
        rows = [ln for ln in f.read().splitlines() if ln]
    dom = sim._domain
    y_last = float(rows[-1].split()[1])
    assert y_last >= dom[1] + 0.9 * (dom[4] - dom[1]) and all(float(r.split()[1]) < dom[1] + 0.9 * (dom[4] - dom[1]) for r in rows[:-1])
    sim.close()


def test_contact_loop_replays_the_reference_draws(golden_small):
    """monte_carlo::init, then 12 x { step ; save_metrics ; repopulate_contacts } with the reference's own rand() stream,
    split per exciton (ids in order of birth; tests/golden/make_golden_contacts.py): every surviving exciton ends on the
    reference's site with the reference's position, free-flight time and heading, bit for bit, and the integer bins
    reproduce the reference's population_profile.dat and region_current.dat."""
    import os
    g = golden_small
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_forster_contacts_replay.npz"))
    e = Engine(g.mc)
    e.set_mesh(g.pos_nm, g.orient)
    e.init_replay(int(z["c1_pop"]), int(z["c2_pop"]), z["draw_off"], z["draws"], z["draw_logs"])
    assert e.number_of_particles() == int(z["p0"])
    n_it, dt = int(z["iterations"]), float(z["dt"])
    pop_a, cur_a = e.step(dt, 5)                 # several iterations per launch, and a second call
    pop_b, cur_b = e.step(dt, n_it - 5)
    pop, cur = np.concatenate([pop_a, pop_b]), np.concatenate([cur_a, cur_b])
    p, ids = e.particles(), e.gids().astype(np.int64)
    order, ref_order = np.argsort(ids), np.argsort(z["ids"])
    assert np.array_equal(ids[order], z["ids"][ref_order])                       # the same excitons survive
    assert np.array_equal(p["site"][order], z["site"][ref_order])
    assert np.array_equal(p["pos"][:, order], z["pos"][:, ref_order])
    assert np.array_equal(p["ff"][order], z["ff"][ref_order])
    assert np.array_equal(p["heading"][order].astype(np.int32), z["heading"][ref_order])
    n_seg = e.number_of_segments()
    area, dom = e.area(), e.domain()
    dy = (dom[4] - dom[1]) / n_seg
    rows = lambda text, n: np.array([[float(v) for v in ln.split(",")] for ln in text.splitlines() if ln and ln[0] in "+-"]).reshape(-1, n + 1)
    pop_ref, cur_ref = rows(bytes(z["pop_file"]).decode(), n_seg), rows(bytes(z["curr_file"]).decode(), n_seg - 1)
    assert np.allclose(pop / (area * dy), pop_ref[:, 1:], rtol=2e-6, atol=0)    # the files carry 7 significant digits
    assert np.allclose(cur / ((area[:-1] + area[1:]) / 2 * dt), cur_ref[:, 1:], rtol=2e-6, atol=0)
    with pytest.raises(CntmcError) as ei:                                       # the lists end with the recorded run
        e.step(dt, 1)
    assert ei.value.code == -4
