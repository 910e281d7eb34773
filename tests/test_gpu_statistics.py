"""Physics-level checks on the GPU: an analytic diffusion coefficient (pure continuous-time random walk on a lattice of
link-less sites, SURVEY.md §8c pin 6), the telescoping property of the displacement accumulator, and the 3-sigma
ensemble agreement of diffusion coefficients with the reference-faithful oracle for independent seeds."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from oracle import t1 as T1m
from conftest import base_mc

pytestmark = pytest.mark.gpu


def constant_table(rate):
    g = np.array([0.0, 1.0])
    return dict(theta=g, z=g, a1=g, a2=g, rates=np.full((2, 2, 2, 2), rate))


def test_lattice_random_walk_has_the_analytic_diffusion_coefficient():
    n, a, r = 41, 5.0, 1e12                      # 41^3 sites, pitch 5 nm, 6 neighbours inside 6 nm, rate 1e12 /s each
    pos, ori = film.lattice_film(n, a)
    mc = base_mc(**{"max hopping radius [m]": 6e-9, "number of sections for injection region": 41,
                    "number of particles for kubo simulation": 40000})
    e = Engine(mc)
    e.set_mesh(pos, ori)
    t = constant_table(r)
    e.set_rate_table(t["theta"], t["z"], t["a1"], t["a2"], t["rates"])
    e.kubo_init()
    s = e.sites()
    inner = (s["pos"] > 10e-9).all(axis=0) & (s["pos"] < 190e-9).all(axis=0)
    assert np.allclose(s["max_rate"][inner], 6 * r, rtol=1e-12)    # Gamma_i = sum of the six rates
    assert len(e.inject()) == 1                                    # the central site
    e.kubo_create_particles(40000, seed=2)
    dt, nsteps = 1e-13, 120
    msd = e.kubo_step(dt, nsteps)
    assert e.reinjections() < 40   # the walk stays inside the removal box (re-injection would not bias delta_pos anyway)
    tt = dt * np.arange(1, nsteps + 1)
    D = np.array([np.polyfit(tt, msd[:, c], 1)[0] / 2 for c in range(3)])
    expect = r * (a * 1e-9) ** 2                 # D_x = (1/2) * sum_j rate * dx_j^2 = r a^2
    # statistical error of the slope: each exciton makes ~70 hops; relative sigma of <x^2> ~ sqrt(2/40000) ~ 0.7 %
    assert np.allclose(D, expect, rtol=0.03), (D, expect)
    assert e.hops() == pytest.approx(40000 * 6 * r * dt * nsteps, rel=0.01)


def test_displacement_telescopes_when_nothing_is_reinjected():
    """delta_pos = pos - start for every exciton as long as no re-injection happened (SURVEY.md §8c pin 7)."""
    pos, ori = film.film(**film.CONFIG_FILMS["C1"])
    e = Engine(base_mc())
    e.set_mesh(pos, ori)
    e.kubo_init()
    e.kubo_create_particles(20000, seed=6)
    p0 = e.particles()
    e.kubo_step(2e-14, 40, want_msd=False)
    p1 = e.particles()
    if e.reinjections() == 0:
        assert np.allclose(p1["delta"], p1["pos"] - p0["pos"], rtol=0, atol=1e-20)
    else:  # excitons that never left the removal box still telescope
        moved = np.abs(p1["delta"] - (p1["pos"] - p0["pos"])).max(axis=0) > 1e-20
        assert moved.sum() <= e.reinjections()


def test_diffusion_coefficient_agrees_with_the_oracle_within_three_sigma():
    """Independent seeds on both sides, E ensembles each: |mean_ref - mean_new| <= 3 sqrt(s_ref^2/E + s_new^2/E)."""
    pos, ori = film.film(NT=150, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
    mc = base_mc()
    E, P, dt, nsteps = 16, 400, 1e-13, 60
    tt = dt * np.arange(1, nsteps + 1)
    half = slice(nsteps // 2, None)

    def slopes(msd):
        return np.array([np.polyfit(tt[half], msd[half, c], 1)[0] / 2 for c in range(3)])

    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.kubo_init()
    d_new = []
    for k in range(E):
        e.kubo_create_particles(P, seed=1000 + k)
        d_new.append(slopes(e.kubo_step(dt, nsteps)))
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    t.set_memo(True)
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    d_ref = []
    for k in range(E):
        t.draws_glibc()                 # the reference's own generator, different srand() seeds
        libc.srandom(5000 + k)
        t.create_particles(P)
        d_ref.append(slopes(t.kubo_step(dt, nsteps)))
    d_new, d_ref = np.array(d_new), np.array(d_ref)
    sigma = np.sqrt(d_new.var(axis=0, ddof=1) / E + d_ref.var(axis=0, ddof=1) / E)
    diff = np.abs(d_new.mean(axis=0) - d_ref.mean(axis=0))
    assert (diff <= 3 * sigma).all(), (diff, sigma, d_new.mean(axis=0), d_ref.mean(axis=0))
