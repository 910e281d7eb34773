"""BASELINE config 2 at full size (1e5 sites, 1e6 excitons) through the C ABI: the oracle cannot run this in seconds,
so the checks are size-independent properties of the path plus an oracle spot check on a sample of
excitons (streams are keyed by global id, so any sub-population can be re-run alone)."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from oracle import t1 as T1m
from conftest import base_mc

pytestmark = pytest.mark.gpu

P, DT, STEPS = 1_000_000, 1e-13, 48


@pytest.fixture(scope="module")
def c2():
    pos, ori = film.film(**film.CONFIG_FILMS["C2"])
    mc = base_mc(**{"number of particles for kubo simulation": P})
    return mc, pos, ori


def run(c2, opts, first=0, count=P, steps=(STEPS,), trace=0):
    mc, pos, ori = c2
    e = Engine(mc)
    e.set_mesh(pos, ori)
    for k, v in opts.items():
        e.set_option(k, v)
    e.kubo_init()
    e.kubo_create_particles(count, seed=1, first_global_id=first)
    p0 = e.particles()
    msd = np.concatenate([e.kubo_step(DT, n) for n in steps])
    return e, p0, e.particles(), msd


def test_full_size_properties(c2):
    e, p0, p1, msd = run(c2, dict(chunk_steps=64, deep_thr=16))   # with the trap solver; the comparison run below without
    # every draw is accounted for: creation takes 3 (site, free flight, heading), an event 2, a re-injection 1
    assert int(p1["ndraw"].astype(np.int64).sum()) == 3 * P + 2 * e.hops() + e.reinjections()
    assert e.hops() > 20 * P * 0.5
    # displacement telescopes for every exciton that was not re-injected (SURVEY.md §8c pin 7)
    off = np.abs(p1["delta"] - (p1["pos"] - p0["pos"])).max(axis=0) > 1e-18
    assert off.sum() <= e.reinjections()
    # the ensemble rows are the mean of squares of what the excitons carry at the end
    assert np.allclose(msd[-1], (p1["delta"] ** 2).mean(axis=1), rtol=1e-12, atol=0)
    # early on the motion is ballistic along the tubes: the MSD grows at every step (later the excitons turn around at
    # the tube ends -- 100 sites x 5 nm against 20 nm of flight per step -- and it may shrink again)
    assert np.all(np.diff(msd.sum(axis=1)[:12]) > 0) and np.all(msd > 0)
    # launch size, scheduling options and splitting the call change nothing, bit for bit
    e2, _, q1, msd2 = run(c2, dict(chunk_steps=16, hot_pct=0, occupancy=6, deep_thr=0), steps=(5, STEPS - 5))
    assert all(np.array_equal(p1[k], q1[k]) for k in p1)
    assert np.array_equal(msd, msd2) and e2.hops() == e.hops()
    # so does running the trap kernel beside the lane kernel, with excitons handed over while both run
    e3, _, r1, msd3 = run(c2, dict(chunk_steps=64, deep_thr=8, deep_group=1, deep_overlap=1, trap_burst=4))
    assert all(np.array_equal(p1[k], r1[k]) for k in p1)
    assert np.array_equal(msd, msd3) and e3.hops() == e.hops()
    # a shard of the population run alone follows the same trajectories (what multi-GPU sharding relies on)
    first, count = 700_000, 4096
    _, _, s1, _ = run(c2, dict(chunk_steps=7), first=first, count=count)
    assert all(np.array_equal(s1[k], p1[k][..., first:first + count]) for k in s1)
    # and the oracle agrees on that shard: same sites, same number of draws, positions within libm log's last ulp
    mc, pos, ori = c2
    t = T1m.T1()
    t.draws_philox(1)
    t.set_memo(True)
    t.kubo_init(mc, pos, ori)
    n_o = 256
    t.create_particles(n_o, first_global_id=first)
    t.kubo_step(DT, STEPS, want_msd=False)
    pt = t.particles()
    assert np.array_equal(pt["site"], s1["site"][:n_o]) and np.array_equal(pt["heading"].astype(np.uint8), s1["heading"][:n_o])
    assert np.allclose(pt["pos"], s1["pos"][:, :n_o], rtol=1e-9, atol=1e-18) and np.allclose(pt["delta"], s1["delta"][:, :n_o], rtol=1e-9, atol=1e-18)


def test_host_resident_population_in_slices(c2):
    """cntmc_kubo_step_host_state steps a big uploaded population in slices on their own streams (copies overlap kernels):
    per-exciton results are the bits of the resident run, the ensemble rows agree to the last bits' summation order."""
    mc, pos, ori = c2
    n = 400_000
    runs = []
    for slices in (1, 4, 3):
        e = Engine(mc)
        e.set_mesh(pos, ori)
        e.set_option("host_slices", slices)
        e.kubo_init()
        e.kubo_create_particles(n, seed=5, first_global_id=1000)
        state = e.particles()
        msd = np.concatenate([e.kubo_step_host_state(DT, 40, state), e.kubo_step_host_state(DT, 25, state)])
        runs.append((state, msd, e.hops(), e.reinjections(), e.time()))
    res = Engine(mc)
    res.set_mesh(pos, ori)
    res.kubo_init()
    res.kubo_create_particles(n, seed=5, first_global_id=1000)
    msd_r = np.concatenate([res.kubo_step(DT, 40), res.kubo_step(DT, 25)])
    pr = res.particles()
    for state, msd, hops, reinj, t in runs:
        assert all(np.array_equal(state[k], pr[k]) for k in pr)
        assert np.allclose(msd, msd_r, rtol=1e-12, atol=0)
        assert hops == res.hops() and reinj == res.reinjections() and t == res.time()
