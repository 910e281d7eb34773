"""Generate tests/golden/*.npz from the reference's own code (oracle/_ref/libt0.so, built from /root/reference).

Run in the dev container only (it needs /root/reference to have been compiled by `make -C oracle t0`):

    python tests/golden/make_golden.py

Each case stores: the synthetic film, the "exciton monte carlo" JSON block, and what the reference computed from them
with srand(seed), single thread: rate table, post-trim site list, domain, removal box, injection list, every site's
neighbour list with cumulative rates (scatterer::find_neighbors), the rand() draws each exciton consumed, the host
libm's log(r/RAND_MAX) of each draw, exciton states after creation and after the run, and the MSD rows.  It also runs
the verbatim reference program (oracle/_ref/cnt_mc_ref = src/main.cpp) and stores its output file.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cnt_film_monte_carlo_b200 import film  # noqa: E402
from oracle import t0 as T0m  # noqa: E402
from oracle import t1 as T1m  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "small_forster": dict(
        film=dict(NT=20, NP=30, a=5.0, LX=60.0, LY=40.0, seed=7),
        mc={
            "rate type": "forster",
            "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
            "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21],
            "temperature [kelvin]": 300, "max hopping radius [m]": 20e-9, "number of segments": 10,
            "trim limits": {"xlim": [-50e-9, 110e-9], "ylim": [0, 40e-9], "zlim": [-50e-9, 110e-9]},
            "exciton velocity [m/s]": 2e5, "monte carlo time step": 1e-14,
            "number of sections for injection region": 5,
            "maximum time for kubo simulation [seconds]": 3e-12,
            "number of particles for kubo simulation": 50,
        },
        seed=100, nsteps=300, contact_steps=80, track_dt=2e-12),
    "wong_trimmed": dict(
        film=dict(NT=30, NP=40, a=4.0, LX=80.0, LY=50.0, seed=11),
        mc={
            "rate type": "wong",
            "zshift [m]": [1.0e-9, 8e-9, 8], "axis shift 1 [m]": [-8e-9, 8e-9, 9],
            "axis shift 2 [m]": [-8e-9, 8e-9, 9], "theta [degrees]": [0, 180, 13],
            "temperature [kelvin]": 300, "max hopping radius [m]": 15e-9, "number of segments": 6,
            "trim limits": {"xlim": [-20e-9, 95e-9], "ylim": [5e-9, 45e-9], "zlim": [-10e-9, 90e-9]},
            "exciton velocity [m/s]": 1.5e5, "monte carlo time step": 2e-13,
            "number of sections for injection region": 3,
            "maximum time for kubo simulation [seconds]": 4e-11,
            "number of particles for kubo simulation": 40,
        },
        seed=2024, nsteps=200),
}


def make(name, case):
    pos, ori = film.film(**case["film"])
    with tempfile.TemporaryDirectory() as tmp:
        mesh, out = os.path.join(tmp, "mesh"), os.path.join(tmp, "out")
        film.write_mesh(mesh, pos, ori)
        pos, ori = film.read_mesh(mesh)  # what every reader sees
        mc = dict(case["mc"])
        mc.update({"mesh input directory": mesh, "output directory": out, "keep old results": False})
        jpath = os.path.join(tmp, "input.json")
        with open(jpath, "w") as f:
            json.dump({"exciton monte carlo": mc}, f)
        P, dt, nsteps, seed = mc["number of particles for kubo simulation"], mc["monte carlo time step"], case["nsteps"], case["seed"]

        t = T0m.T0()
        t.open(jpath, seed)
        tab, sites = t.table(), t.sites()
        row_ptr, nbr, cum = t.csr()
        g = dict(pos_nm=pos, orient=ori, nsteps=nsteps, seed=seed, dt=dt,
                 **{"table_" + k: v for k, v in tab.items()}, **{"site_" + k: v for k, v in sites.items()},
                 domain=t.domain(), removal=t.removal_domain(), inject=t.inject(), row_ptr=row_ptr, nbr=nbr, cum=cum)
        t.log_draws(True)
        t.srand(seed)
        t.create_particles_logged(P)
        g.update({"p0_" + k: v for k, v in t.particles().items()})
        g["msd"] = t.kubo_step_logged(dt, nsteps)
        g.update({"p1_" + k: v for k, v in t.particles().items()})
        off, flat = t.draws(P)
        g["draw_off"], g["draws"], g["draw_logs"] = off, flat, T1m.log_ratios(flat)
        g["time"] = t.time()
        # the same run through the reference's own kubo_create_particles / kubo_step (no attribution) must agree
        t.open(jpath, seed)
        t.srand(seed)
        t.create_particles_verbatim()
        msd_v = t.kubo_step_verbatim(dt, nsteps)
        assert np.array_equal(msd_v, g["msd"]), "logged loop differs from the reference's kubo_step"
        pv = t.particles()
        assert all(np.array_equal(pv[k], g["p1_" + k]) for k in pv)
        t.close()
        # the verbatim reference program, for its output file
        mc["maximum time for kubo simulation [seconds]"] = dt * (nsteps - 0.5)
        with open(jpath, "w") as f:
            json.dump({"exciton monte carlo": mc}, f)
        if seed == 100:  # main.cpp:30 hard-codes srand(100)
            subprocess.check_call([T0m.REF_BINARY, jpath], stdout=subprocess.DEVNULL, env=dict(os.environ, OMP_NUM_THREADS="1"))
            with open(os.path.join(out, "particle_dispalcement.avg.squared.dat")) as f:
                g["ref_program_output"] = np.frombuffer(f.read().encode(), dtype=np.uint8)
        # ---- contact mode (monte_carlo::init + the loop at main.cpp:98-106), reference code end to end ----------------
        if case.get("contact_steps"):
            mc2 = dict(case["mc"])
            mc2.update({"mesh input directory": mesh, "output directory": os.path.join(tmp, "out_contacts"), "keep old results": False})
            with open(jpath, "w") as f:
                json.dump({"exciton monte carlo": mc2}, f)
            t.open_contacts(jpath, seed)
            n_seg = int(mc2["number of segments"])
            g["contact_area"] = t.area(n_seg)
            g["contact_c1"], g["contact_c2"] = t.contact_sites(1), t.contact_sites(2)
            g["contact_p0_site"] = t.particles()["site"]
            npart = [t.L.t0_num_particles()]
            for _ in range(case["contact_steps"]):
                t.contact_iteration(dt)
                npart.append(t.L.t0_num_particles())
            g["contact_num_particles"] = np.array(npart)
            # monte_carlo::track_particle (monte_carlo.h:786-818) with its draws logged; writes particle_path.3.dat
            t.L.t0_clear_draws()
            t.log_draws(True)
            t.track_particle(case["track_dt"], 3, 0)
            _, tflat = t.draws(1)
            g["track_dt"], g["track_draws"], g["track_logs"] = case["track_dt"], tflat, T1m.log_ratios(tflat)
            t.close()
            for name_, key in (("population_profile.dat", "contact_pop_file"), ("region_current.dat", "contact_curr_file"),
                               ("scatterer_statistics.dat", "contact_stat_file"), ("particle_path.3.dat", "track_file")):
                with open(os.path.join(tmp, "out_contacts", name_)) as f:
                    g[key] = np.frombuffer(f.read().encode(), dtype=np.uint8)
    mc_clean = dict(case["mc"])
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump({"exciton monte carlo": mc_clean}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **g)
    print(name, "sites", len(sites["left"]), "nnz", len(nbr), "draws", len(flat), "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    if not T0m.available():
        sys.exit("oracle/_ref/libt0.so missing: run `make -C oracle t0` first")
    for name, case in CASES.items():
        make(name, case)
