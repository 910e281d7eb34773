"""The T1 oracle (plain-C restatement) against the golden vectors computed by the reference's own code, and -- when
oracle/_ref/libt0.so exists (dev container) -- against that code live on a larger film."""
import json
import os

import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from oracle import t0 as T0m
from oracle import t1 as T1m
from conftest import base_mc

STATE_KEYS = ("site", "pos", "delta", "ff", "heading")


def setup_t1(g):
    s = T1m.T1()
    s.kubo_init(g.mc, g.pos_nm, g.orient)
    return s


def test_setup_matches_reference(golden):
    s = setup_t1(golden)
    tab = T1m.forster_table(golden.mc)
    for k, v in golden.group("table_").items():
        assert np.array_equal(tab[k], v), k
    sites = s.sites()
    for k, v in golden.group("site_").items():
        assert np.array_equal(sites[k], v), k
    assert np.array_equal(s.domain(), golden.z["domain"])
    assert np.array_equal(s.removal_domain(), golden.z["removal"])
    assert np.array_equal(s.inject(), golden.z["inject"])
    assert sites["left"].size < golden.pos_nm[0].size  # the trim limits really removed sites


def test_neighbour_lists_bit_exact(golden):
    row_ptr, nbr, cum = setup_t1(golden).csr()
    assert np.array_equal(row_ptr, golden.z["row_ptr"])
    assert np.array_equal(nbr, golden.z["nbr"])
    assert np.array_equal(cum, golden.z["cum"])  # bit for bit: same enumeration order, same sequential sum


def test_replayed_draws_reproduce_reference_run(golden):
    s = setup_t1(golden)
    s.draws_replay(golden.z["draw_off"], golden.z["draws"])
    s.log_draws(True)
    s.create_particles(golden.P)
    p0 = s.particles()
    for k, v in golden.group("p0_").items():
        assert np.array_equal(p0[k], v), k
    msd = s.kubo_step(golden.dt, golden.nsteps)
    assert np.array_equal(msd, golden.z["msd"])
    p1 = s.particles()
    for k, v in golden.group("p1_").items():
        assert np.array_equal(p1[k], v), k
    assert not s.replay_exhausted()
    off, flat = s.logged_draws(golden.P)
    assert np.array_equal(off, golden.z["draw_off"]) and np.array_equal(flat, golden.z["draws"])  # consumed exactly
    assert s.time() == float(golden.z["time"])
    assert np.array_equal(T1m.log_ratios(flat), golden.z["draw_logs"])


def test_glibc_stream_reproduces_reference_run(golden):
    import ctypes
    s = setup_t1(golden)
    s.draws_glibc()
    ctypes.CDLL("libc.so.6").srandom(golden.seed)
    s.create_particles(golden.P)
    msd = s.kubo_step(golden.dt, golden.nsteps)
    assert np.array_equal(msd, golden.z["msd"])
    p1 = s.particles()
    for k, v in golden.group("p1_").items():
        assert np.array_equal(p1[k], v), k


def test_reference_program_output_file(golden_small):
    """particle_dispalcement.avg.squared.dat written by the verbatim reference program = our MSD rows, 7 digits."""
    text = bytes(golden_small.z["ref_program_output"]).decode()
    lines = text.splitlines()
    assert lines[0].startswith("# this file contains the average of dx^2")
    assert lines[1] == "# number of particles: %d" % golden_small.P and lines[3] == "time,x,y,z"
    rows = np.array([[float(v) for v in ln.split(",")] for ln in lines[4:]])
    n = min(len(rows), golden_small.nsteps)
    assert n >= golden_small.nsteps - 1
    assert np.allclose(rows[:n, 1:], golden_small.z["msd"][:n], rtol=1e-6)
    assert lines[4].startswith("+1.000000e-14,+")  # showpos + scientific, 6 digits (monte_carlo.cpp:387, 408)


def test_memoised_rows_change_nothing(golden_small):
    a, b = setup_t1(golden_small), setup_t1(golden_small)
    b.set_memo(True)
    for s in (a, b):
        s.draws_philox(5)
        s.create_particles(30)
        s.kubo_step(golden_small.dt, 100, want_msd=False)
    pa, pb = a.particles(), b.particles()
    assert all(np.array_equal(pa[k], pb[k]) for k in pa)


@pytest.mark.skipif(not T0m.available(), reason="oracle/_ref/libt0.so only exists where /root/reference was compiled")
def test_t1_equals_reference_code_live(tmp_path):
    """A film ~10x the golden ones, every site's neighbour list and a 2000-step run, T1 vs the reference TUs."""
    pos, ori = film.film(NT=60, NP=50, a=5.0, LX=150.0, LY=60.0, seed=99)
    mesh = str(tmp_path / "mesh")
    film.write_mesh(mesh, pos, ori)
    pos, ori = film.read_mesh(mesh)
    mc = base_mc(**{"mesh input directory": mesh, "output directory": str(tmp_path / "out"), "keep old results": False,
                    "trim limits": {"xlim": [-30e-9, 170e-9], "ylim": [2e-9, 58e-9], "zlim": [-30e-9, 170e-9]}})
    jpath = str(tmp_path / "input.json")
    with open(jpath, "w") as f:
        json.dump({"exciton monte carlo": mc}, f)
    t0 = T0m.T0()
    t0.open(jpath, 100)
    t1 = T1m.T1()
    t1.kubo_init(mc, pos, ori)
    s0, s1 = t0.sites(), t1.sites()
    assert all(np.array_equal(s0[k], s1[k]) for k in s0)
    for a, b in zip(t0.csr(threads=os.cpu_count()), t1.csr()):
        assert np.array_equal(a, b)
    t0.srand(321)
    t0.create_particles_verbatim()
    m0 = t0.kubo_step_verbatim(2e-14, 2000)
    t1.draws_glibc()
    t0.srand(321)
    t1.create_particles(int(mc["number of particles for kubo simulation"]))
    m1 = t1.kubo_step(2e-14, 2000)
    assert np.array_equal(m0, m1)
    p0, p1 = t0.particles(), t1.particles()
    assert all(np.array_equal(p0[k], p1[k]) for k in STATE_KEYS)
    t0.close()
