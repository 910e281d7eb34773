"""The reference-shaped host sides (C++ driver cntmc_main over cpp/monte_carlo.hpp, and the Python class
cnt_film_monte_carlo_b200.monte_carlo) end to end on the GPU: same CLI, same input.json, same output files."""
import json
import os
import subprocess

import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import build, film
from cnt_film_monte_carlo_b200.engine import Engine
from cnt_film_monte_carlo_b200.monte_carlo import monte_carlo
from oracle import t1 as T1m

pytestmark = pytest.mark.gpu


def write_case(tmp_path, g, out_name, **over):
    mesh = str(tmp_path / "mesh")
    film.write_mesh(mesh, g.pos_nm, g.orient)
    mc = dict(g.mc)
    mc.update({"mesh input directory": mesh, "output directory": str(tmp_path / out_name), "keep old results": False})
    mc.update(over)
    path = str(tmp_path / (out_name + ".json"))
    with open(path, "w") as f:
        json.dump({"exciton monte carlo": mc, "cnts": {"comment": "unused for forster"}}, f)
    return path, mc


def read_rows(path, skip):
    with open(path) as f:
        lines = f.read().splitlines()
    return lines[:skip], np.array([[float(v) for v in ln.split(",")] for ln in lines[skip:] if ln])


def test_cpp_driver_green_kubo_run_matches_oracle(tmp_path, golden_small):
    g = golden_small
    nsteps = 120
    path, mc = write_case(tmp_path, g, "out_cpp", **{"maximum time for kubo simulation [seconds]": g.dt * (nsteps - 0.5)})
    r = subprocess.run([build.build_driver(), path, "--steps-per-call", "50", "--seed", "100"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Green-Kubo simulation finished!" in r.stdout
    head, rows = read_rows(os.path.join(mc["output directory"], "particle_dispalcement.avg.squared.dat"), 4)
    assert head[1] == "# number of particles: %d" % g.P and head[3] == "time,x,y,z"
    assert len(rows) == nsteps
    t = T1m.T1()
    t.kubo_init(g.mc, g.pos_nm, g.orient)
    t.draws_philox(100)
    t.create_particles(g.P)
    msd = t.kubo_step(g.dt, nsteps)
    assert np.allclose(rows[:, 1:], msd, rtol=2e-6) and np.allclose(rows[:, 0], g.dt * np.arange(1, nsteps + 1), rtol=1e-6)
    saved = json.load(open(os.path.join(mc["output directory"], "input.json")))
    assert saved["rate type"] == "forster" and saved["max hopping radius [m]"] == 20e-9
    # same run through the Python mirror of the class: identical file
    path2, mc2 = write_case(tmp_path, g, "out_py", **{"maximum time for kubo simulation [seconds]": g.dt * (nsteps - 0.5)})
    sim = monte_carlo(mc2, seed=100, quiet=True)
    sim.kubo_init()
    sim.save_json_properties()
    sim.kubo_create_particles()
    while sim.time() < sim.kubo_max_time():
        sim.kubo_step(g.dt)
        sim.kubo_save_avg_dispalcement_squared()
    sim.close()
    a = open(os.path.join(mc["output directory"], "particle_dispalcement.avg.squared.dat")).read()
    b = open(os.path.join(mc2["output directory"], "particle_dispalcement.avg.squared.dat")).read()
    assert a == b


def test_cpp_driver_contact_run_writes_reference_format(tmp_path, golden_small):
    g = golden_small
    path, mc = write_case(tmp_path, g, "out_contacts")
    r = subprocess.run([build.build_driver(), path, "--contacts", "30", "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ref_pop = bytes(g.z["contact_pop_file"]).decode().splitlines()
    ref_cur = bytes(g.z["contact_curr_file"]).decode().splitlines()
    pop = open(os.path.join(mc["output directory"], "population_profile.dat")).read().splitlines()
    cur = open(os.path.join(mc["output directory"], "region_current.dat")).read().splitlines()
    assert pop[:7] == ref_pop[:7] and cur[:5] == ref_cur[:5]     # header blocks: areas, dy, positions, column names
    assert len(pop) == 7 + 30 and len(cur) == 5 + 30
    t = T1m.T1()
    t.draws_philox(5)
    t.contacts_init(g.mc, g.pos_nm, g.orient)
    area, dom = t.area(), t.domain()
    dy = (dom[4] - dom[1]) / 10
    for s in range(30):
        p, c = t.contact_iteration(g.dt)
        got = np.array([float(v) for v in pop[7 + s].split(",")])
        assert np.allclose(got[1:], p / (area * dy), rtol=2e-6)
        gotc = np.array([float(v) for v in cur[5 + s].split(",")])
        assert np.allclose(gotc[1:], c / ((area[:-1] + area[1:]) / 2 * g.dt), rtol=2e-6)


def test_driver_reports_errors_like_the_reference(tmp_path):
    bad = tmp_path / "bad.json"
    bad.write_text('{"cnts": {}}')
    r = subprocess.run([build.build_driver(), str(bad)], capture_output=True, text=True)
    assert r.returncode != 0 and 'does not contain "exciton monte carlo"' in r.stderr


def test_cpp_driver_writes_individual_displacements(tmp_path, golden_small):
    """monte_carlo::kubo_save_individual_particle_dispalcements (monte_carlo.cpp:345-380) in the C++ shim: three files, a
    header line `time,+0,+1,...` (showpos), one row per call with the exciton's accumulated displacement -- and the same
    text from the Python mirror of the class."""
    g = golden_small
    nsteps = 60
    path, mc = write_case(tmp_path, g, "out_disp", **{"maximum time for kubo simulation [seconds]": g.dt * (nsteps - 0.5)})
    r = subprocess.run([build.build_driver(), path, "--steps-per-call", "20", "--seed", "100", "--displacements", "1"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    e = Engine(g.mc)
    e.set_mesh(g.pos_nm, g.orient)
    e.kubo_init()
    e.kubo_create_particles(g.P, seed=100)
    want = []
    for _ in range(3):
        e.kubo_step(g.dt, 20, want_msd=False)
        want.append(e.particles()["delta"].copy())
    for c, ax in enumerate("xyz"):
        with open(os.path.join(mc["output directory"], "particle_dispalcement.%s.dat" % ax)) as f:
            lines = f.read().splitlines()
        assert lines[0] == "time" + "".join(",%+d" % i for i in range(g.P))
        assert len(lines) == 4
        for k in range(3):
            vals = [float(v) for v in lines[1 + k].split(",")]
            assert abs(vals[0] - g.dt * 20 * (k + 1)) < 1e-20 and lines[1 + k].startswith("+")
            assert np.allclose(vals[1:], want[k][c], rtol=2e-6, atol=1e-30)
