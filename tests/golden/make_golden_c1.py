"""Generate tests/golden/c1_input_json.npz: BASELINE config 1 = the reference's own input.json (trim limits, dt = 1e-15 s,
hopping radius, grids; input.json:46-67 verbatim except "rate type": the closed-form table instead of "davoody") on the
synthetic stand-in for its missing mesh (film.CONFIG_FILMS["C1"]), run by the reference's own code (oracle/_ref/libt0.so).

    python tests/golden/make_golden_c1.py        # dev container only: needs /root/reference compiled by `make -C oracle t0`

The film is deterministic (seed 1234) and is not stored.  Stored: the post-trim chain links and site count, sampled site
positions, domain / removal box / injection list, every site's degree, the neighbour lists of every 41st site in full and
wrap-around checksums of all of them, and for the first P_LOG excitons of the reference's srand(100) run their rand() draws,
states after creation and after NSTEPS steps of 1e-15 s, plus every 50th MSD row.
"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cnt_film_monte_carlo_b200 import film  # noqa: E402
from oracle import t0 as T0m  # noqa: E402
from oracle import t1 as T1m  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P_LOG, NSTEPS, ROW_STRIDE, MSD_STRIDE = 200, 20000, 41, 50
MC = {
    "rate type": "forster",
    "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
    "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21],
    "temperature [kelvin]": 300, "max hopping radius [m]": 20e-9, "number of segments": 10,
    "trim limits": {"xlim": [-1e-6, 1e-6], "ylim": [0, 1e-7], "zlim": [-1e-6, 1e-6]},
    "exciton velocity [m/s]": 2e5, "monte carlo time step": 1e-15,
    "number of sections for injection region": 5,
    "maximum time for kubo simulation [seconds]": 1e-8,
    "number of particles for kubo simulation": P_LOG,
}


def checksums(row_ptr, nbr, cum):
    k = np.arange(len(nbr), dtype=np.uint64)
    return np.array([np.sum(nbr.astype(np.uint64) * (k % np.uint64(1000003) + np.uint64(1)), dtype=np.uint64),
                     np.sum(cum.view(np.uint64) ^ (k * np.uint64(0x9E3779B97F4A7C15)), dtype=np.uint64)], np.uint64)


def main():
    pos, ori = film.film(**film.CONFIG_FILMS["C1"])
    with tempfile.TemporaryDirectory() as tmp:
        mesh = os.path.join(tmp, "mesh")
        film.write_mesh(mesh, pos, ori)
        p2, o2 = film.read_mesh(mesh)
        assert np.array_equal(p2, pos) and np.array_equal(o2, ori)  # the text round-trips: tests can feed the film directly
        mc = dict(MC)
        mc.update({"mesh input directory": mesh, "output directory": os.path.join(tmp, "out"), "keep old results": False})
        jpath = os.path.join(tmp, "input.json")
        with open(jpath, "w") as f:
            json.dump({"exciton monte carlo": mc}, f)
        t = T0m.T0()
        t.open(jpath, 100)
        sites = t.sites()
        row_ptr, nbr, cum = t.csr()
        N = len(sites["left"])
        g = dict(n_sites=N, site_left=sites["left"], site_right=sites["right"], site_pos_sample=sites["pos"][:, ::ROW_STRIDE],
                 site_max_rate_sample=sites["max_rate"][::ROW_STRIDE], domain=t.domain(), removal=t.removal_domain(), inject=t.inject(),
                 degree=np.diff(row_ptr).astype(np.uint16), csr_checksums=checksums(row_ptr, nbr, cum), dt=MC["monte carlo time step"],
                 nsteps=NSTEPS)
        rows = np.arange(0, N, ROW_STRIDE)
        g["row_sample_nbr"] = np.concatenate([nbr[row_ptr[i]:row_ptr[i + 1]] for i in rows])
        g["row_sample_cum"] = np.concatenate([cum[row_ptr[i]:row_ptr[i + 1]] for i in rows])
        t.log_draws(True)
        t.srand(100)
        t.create_particles_logged(P_LOG)
        g.update({"p0_" + k: v for k, v in t.particles().items() if k != "old_pos"})
        msd = t.kubo_step_logged(MC["monte carlo time step"], NSTEPS)
        g["msd_sample"] = msd[MSD_STRIDE - 1::MSD_STRIDE]
        g.update({"p1_" + k: v for k, v in t.particles().items() if k != "old_pos"})
        off, flat = t.draws(P_LOG)
        g["draw_off"], g["draws"], g["draw_logs"] = off, flat, T1m.log_ratios(flat)
        t.close()
    with open(os.path.join(HERE, "c1_input_json.json"), "w") as f:
        json.dump({"exciton monte carlo": MC}, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "c1_input_json.npz"), **g)
    print("sites", N, "of", pos.shape[1] * pos.shape[2], "nnz", len(nbr), "draws", len(flat), "bytes", os.path.getsize(os.path.join(HERE, "c1_input_json.npz")))


if __name__ == "__main__":
    main()
