"""One simulation on several GPUs through the product's own entry points (cntmc_multi_*, include/cntmc.h): excitons split
by global id over the GPUs of the box, one ncclAllReduce per call, NCCL bound at run time inside libcntmc.so.
With one visible GPU the single-device cases run (NCCL with one rank); the two-GPU cases need `gpurun --gpus 2`."""
import json
import os
import subprocess

import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import build, film
from cnt_film_monte_carlo_b200.engine import Engine, MultiEngine
from conftest import base_mc

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def film_and_mc():
    pos, ori = film.film(NT=150, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
    return base_mc(), pos, ori


@pytest.mark.parametrize("devices", [[0], [0, 1], [1, 0]])
def test_multi_kubo_equals_single_handle(devices):
    if max(devices) >= n_gpus():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    mc, pos, ori = film_and_mc()
    one = Engine(mc)
    one.set_mesh(pos, ori)
    one.kubo_init()
    one.kubo_create_particles(5001, seed=4)
    msd1 = np.concatenate([one.kubo_step(1e-13, 50), one.kubo_step(1e-13, 37)])
    m = MultiEngine(mc, devices)
    assert m.nccl_version() > 20000
    m.set_mesh(pos, ori)
    m.kubo_init()
    m.kubo_create_particles(5001, seed=4)
    msd = np.concatenate([m.kubo_step(1e-13, 50), m.kubo_step(1e-13, 37)])
    p1, pm = one.particles(), m.particles()
    assert all(np.array_equal(p1[k], pm[k]) for k in p1)             # every trajectory: the same bits for any number of GPUs
    assert np.allclose(msd, msd1, rtol=1e-12, atol=0)                 # ensemble rows: summation order across GPUs differs
    assert m.hops() == one.hops() and m.time() == one.time() and m.number_of_particles() == 5001


@pytest.mark.parametrize("devices", [[0], [0, 1]])
def test_multi_contacts_conserve_and_match_single_gpu_statistics(devices):
    if max(devices) >= n_gpus():
        pytest.skip("needs %d GPUs" % (max(devices) + 1))
    mc, pos, ori = film_and_mc()
    m = MultiEngine(mc, devices)
    m.set_mesh(pos, ori)
    m.init(901, 100, seed=3)
    pop, cur = m.step(1e-14, 30)
    one = Engine(mc)
    one.set_mesh(pos, ori)
    one.init(901, 100, seed=3)
    pop1, cur1 = one.step(1e-14, 30)
    if len(devices) == 1:
        assert np.array_equal(pop, pop1) and np.array_equal(cur, cur1)
    else:  # other streams, same ensemble: the contact slabs hold exactly what repopulate_contacts puts there
        assert np.all(pop[:, 0] > 0.8 * 901) and np.all(pop[:, 0] <= 901 + 100) and pop.sum(axis=1).min() > 0
        assert abs(pop.sum(axis=1).mean() - pop1.sum(axis=1).mean()) < 0.05 * pop1.sum(axis=1).mean()
    assert m.number_of_particles() == pop[-1].sum() - pop[-1, 0] - pop[-1, -1] + 901 + 100


def test_cpp_driver_on_two_gpus_writes_the_same_rows(tmp_path):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    mc, pos, ori = film_and_mc()
    mesh = str(tmp_path / "mesh")
    film.write_mesh(mesh, pos, ori)
    rows = []
    for g in (1, 2):
        cfg = dict(mc)
        cfg.update({"mesh input directory": mesh, "output directory": str(tmp_path / ("out%d" % g)), "keep old results": False,
                    "maximum time for kubo simulation [seconds]": 1e-13 * 79.5, "number of particles for kubo simulation": 3001})
        path = str(tmp_path / ("in%d.json" % g))
        with open(path, "w") as f:
            json.dump({"exciton monte carlo": cfg}, f)
        r = subprocess.run([build.build_driver(), path, "--gpus", str(g), "--steps-per-call", "32", "--displacements", "1"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        with open(os.path.join(cfg["output directory"], "particle_dispalcement.avg.squared.dat")) as f:
            rows.append(np.array([[float(v) for v in ln.split(",")] for ln in f.read().splitlines()[4:] if ln]))
        with open(os.path.join(cfg["output directory"], "particle_dispalcement.x.dat")) as f:
            rows.append(f.read())
    assert np.allclose(rows[0], rows[2], rtol=1e-6) and len(rows[0]) == 80
    assert rows[1] == rows[3]                                        # per-exciton displacements: identical text
