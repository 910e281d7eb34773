"""Contact-driven transport (BASELINE config 5): the T1 oracle against the files written by the reference's own
monte_carlo::init / step / save_metrics / repopulate_contacts (golden fixture), glibc draw stream."""
import ctypes

import numpy as np

from oracle import t1 as T1m

DT = 1e-14


def parse(text, n_cols):
    rows = [ln for ln in text.splitlines() if ln and ln[0] in "+-"]
    return np.array([[float(v) for v in ln.split(",")] for ln in rows]).reshape(len(rows), n_cols + 1)


def test_contact_run_reproduces_reference_files(golden_small):
    g = golden_small
    n_seg = int(g.mc["number of segments"])
    t = T1m.T1()
    t.draws_glibc()
    ctypes.CDLL("libc.so.6").srandom(g.seed)
    t.contacts_init(g.mc, g.pos_nm, g.orient)
    assert np.array_equal(t.area(), g.z["contact_area"])       # incl. the reference's min-or-max quirk (negative areas)
    assert np.array_equal(t.contact_sites(1), g.z["contact_c1"]) and np.array_equal(t.contact_sites(2), g.z["contact_c2"])
    assert np.array_equal(t.particles()["site"], g.z["contact_p0_site"])
    pop_ref = parse(bytes(g.z["contact_pop_file"]).decode(), n_seg)
    cur_ref = parse(bytes(g.z["contact_curr_file"]).decode(), n_seg - 1)
    area = t.area()
    dom = t.domain()
    dy = (dom[4] - dom[1]) / n_seg
    area_if = (area[:-1] + area[1:]) / 2
    nsteps = len(pop_ref)
    assert nsteps == len(g.z["contact_num_particles"]) - 1
    for s in range(nsteps):
        pop, cur = t.contact_iteration(DT)
        assert t.L.t1_num_particles(t.h) == g.z["contact_num_particles"][s + 1]
        # the reference prints count / (area * dy) and net / (area_if * dt) with 7 significant digits
        assert np.allclose(pop / (area * dy), pop_ref[s, 1:], rtol=2e-6, atol=0), s
        assert np.allclose(cur / (area_if * DT), cur_ref[s, 1:], rtol=2e-6, atol=0), s
        assert pop.sum() == g.z["contact_num_particles"][s]     # everybody alive during the step is counted once


def test_philox_contact_run_is_reproducible_and_conserves_excitons(golden_small):
    g = golden_small
    runs = []
    for _ in range(2):
        t = T1m.T1()
        t.draws_philox(11)
        t.contacts_init(g.mc, g.pos_nm, g.orient, c1_pop=300, c2_pop=40)
        hist = [t.contact_iteration(DT) for _ in range(40)]
        runs.append((np.array([h[0] for h in hist]), np.array([h[1] for h in hist])))
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])


def format_path(path):
    """the rows monte_carlo::track_particle writes (monte_carlo.h:806-815): showpos, scientific, 6 digits"""
    return "".join("   %+.6e %+.6e %+.6e\n" % tuple(r) for r in path) + "\n"


def test_track_particle_reproduces_reference_file(golden_small):
    g = golden_small
    flat = g.z["track_draws"]
    t = T1m.T1()
    t.contacts_init(g.mc, g.pos_nm, g.orient)
    t.draws_replay(np.array([0, len(flat)], np.int64), flat)
    path, reached = t.track_particle(float(g.z["track_dt"]), gid=0)
    assert reached and not t.replay_exhausted()
    assert format_path(path) == bytes(g.z["track_file"]).decode()
    extra, _ = t.track_particle(float(g.z["track_dt"]), gid=0, max_steps=len(path) - 1)
    assert len(extra) == len(path) - 1 and np.array_equal(extra, path[:-1])


def test_contact_replay_fixture_is_reproduced_by_the_oracle(golden_small):
    """The per-exciton draw lists of the reference's contact loop (ids in order of birth) replayed by the oracle: the same
    survivors in the same states -- this is the fixture the GPU replays in tests/test_gpu_contacts.py."""
    import os
    g = golden_small
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "small_forster_contacts_replay.npz"))
    t = T1m.T1()
    t.draws_replay(z["draw_off"], z["draws"])
    t.contacts_init(g.mc, g.pos_nm, g.orient, c1_pop=int(z["c1_pop"]), c2_pop=int(z["c2_pop"]))
    for _ in range(int(z["iterations"])):
        t.contact_iteration(float(z["dt"]))
    p = t.particles()
    assert not t.replay_exhausted()
    # T1 keeps the reference's list order (the same swaps), so the lists compare position by position
    assert np.array_equal(p["site"], z["site"]) and np.array_equal(p["pos"], z["pos"]) and np.array_equal(p["ff"], z["ff"])
