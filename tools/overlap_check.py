"""The trap kernel beside the lane kernel (option deep_overlap) against the plain loop on a small film: per-exciton state and
ensemble rows must be bit-identical.  Small enough to run under compute-sanitizer (tools/gpu_r2_sanitize2.sh)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from conftest import base_mc

pos, ori = film.film(NT=150, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
out = []
for opts in (dict(deep_thr=0), dict(deep_thr=8, deep_group=1, deep_overlap=1, trap_burst=4), dict(deep_thr=2, deep_group=1, deep_overlap=1, overlap_trap_blocks=2, chunk_steps=9)):
    e = Engine(base_mc())
    e.set_mesh(pos, ori)
    for k, v in opts.items():
        e.set_option(k, v)
    e.kubo_init()
    e.kubo_create_particles(5000, seed=4)
    m = np.concatenate([e.kubo_step(1e-13, 50), e.kubo_step(1e-13, 37)])
    out.append((e.particles(), m, e.get_option("dbg_deferred")))
for s, m, d in out[1:]:
    assert all(np.array_equal(s[k], out[0][0][k]) for k in s) and np.array_equal(m, out[0][1])
print("overlap ok; deferred in the last launch:", [d for _, _, d in out])
