"""BASELINE config 1 on the GPU: the reference's own input.json (trim limits verbatim, dt = 1e-15 s) on the stand-in
film, against what the reference's code computed (tests/golden/c1_input_json.npz): trim permutation, every neighbour
list, and a replay of the reference's rand() draws over 20 000 steps -- the launch-bound regime."""
import numpy as np
import pytest

import c1_case
from cnt_film_monte_carlo_b200.engine import Engine

pytestmark = pytest.mark.gpu


def test_config_1_setup_and_replay_are_the_reference_bits():
    mc, z, pos, ori = c1_case.load()
    states, rows = [], []
    for opts in (dict(chunk_steps=4096), dict(chunk_steps=64, hot_pct=10), dict(chunk_steps=977, top_entries=0, runs=0, dirs=0)):
        e = Engine(mc)
        e.set_mesh(pos, ori)
        for k, v in opts.items():
            e.set_option(k, v)
        e.kubo_init()
        if not states:
            row_ptr, nbr, cum = e.csr()
            c1_case.check_setup(z, e.sites(), e.domain(), e.removal_domain(), e.inject(), row_ptr, nbr, cum)
            assert e.csr_midpoint_guards() == 0
        e.kubo_create_particles_replay(z["draw_off"], z["draws"], z["draw_logs"])
        p0 = e.particles()
        for k in ("site", "pos", "ff"):
            assert np.array_equal(p0[k], z["p0_" + k]), k
        msd = np.concatenate([e.kubo_step(float(z["dt"]), 12345), e.kubo_step(float(z["dt"]), int(z["nsteps"]) - 12345)])
        p1 = e.particles()
        for k in ("site", "pos", "delta", "ff"):
            assert np.array_equal(p1[k], z["p1_" + k]), k           # bit for bit after 20 000 steps of 1e-15 s
        assert np.array_equal(p1["heading"].astype(np.int32), z["p1_heading"])
        assert np.array_equal(p1["ndraw"], np.diff(z["draw_off"]))  # every recorded draw consumed, none missing
        assert np.allclose(msd[c1_case.MSD_STRIDE - 1::c1_case.MSD_STRIDE], z["msd_sample"], rtol=1e-12, atol=0)
        assert abs(e.time() - 2e-11) < 1e-20
        states.append(p1)
        rows.append(msd)
    assert all(np.array_equal(r, rows[0]) for r in rows[1:])       # launch size and shortcuts: the same bits


def test_config_1_population_of_input_json():
    """The 2000 excitons of input.json on counter-based streams: draw accounting and shard-equals-whole at dt = 1e-15."""
    mc, z, pos, ori = c1_case.load()
    mc = dict(mc)
    mc["number of particles for kubo simulation"] = 2000
    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.set_option("chunk_steps", 4096)
    e.kubo_init()
    e.kubo_create_particles(0, seed=100)
    assert e.number_of_particles() == 2000
    e.kubo_step(1e-15, 30000, want_msd=False)
    p = e.particles()
    assert int(p["ndraw"].astype(np.int64).sum()) == 3 * 2000 + 2 * e.hops() + e.reinjections() and e.hops() > 100000
    s = Engine(mc)
    s.set_mesh(pos, ori)
    s.kubo_init()
    s.kubo_create_particles(300, seed=100, first_global_id=1500)
    s.kubo_step(1e-15, 30000, want_msd=False)
    ps = s.particles()
    assert all(np.array_equal(ps[k], p[k][..., 1500:1800]) for k in ps)
