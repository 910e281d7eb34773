"""Contact mode on the GPU (through the C ABI) against the oracle: populations and interface currents are integers and
must match exactly; so must the surviving population."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
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
