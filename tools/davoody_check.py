"""GPU check + timing of the davoody table builder against the reference units (oracle/_ref/libf1.so).

    python tools/davoody_check.py            # a few tube pairs, random placements, then the input.json-sized table
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import davoody as dv  # noqa: E402
from oracle import f1  # noqa: E402

rng = np.random.default_rng(7)
for (d, a) in [((4, 2, 10), (4, 2, 10)), ((4, 2, 10), (4, 2, 14)), ((5, 3, 4), (5, 3, 4)), ((6, 5, 3), (6, 5, 3)), ((4, 2, 6), (5, 3, 4))]:
    td, ta = dv.Tube(*d), dv.Tube(*a)
    x = dv.Transfer(td, ta)
    rd, ra = f1.RefTube(*d), f1.RefTube(*a)
    n = 6
    z = rng.uniform(1.5e-9, 10e-9, n)
    s1 = rng.uniform(-10e-9, 10e-9, n)
    s2 = rng.uniform(-10e-9, 10e-9, n)
    th = rng.uniform(0, 3.141592, n)
    ours = x.first_order(z, s1, s2, th)
    t0 = time.time()
    ref = np.array([f1.first_order(rd, ra, *g) for g in zip(z, s1, s2, th)])
    t_ref = (time.time() - t0) / n
    rel = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-300)
    print(json.dumps({"donor": d, "acceptor": a, "info": x.info(), "max_rel": float(rel.max()) if ref.any() else 0.0,
                      "bit_equal": int((ours == ref).sum()), "of": n, "ref_s_per_placement": t_ref, "rates": ours[:3].tolist()}))

# the shipped input.json: (4,2) x 10 cells, 21 x 11 x 11 x 11 placements
cfg = {"cnts": {"directory": "x", "comment": "y", "1": {"chirality": [4, 2], "length": [10, "cnt unit cells"]}},
       "exciton monte carlo": {"zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
                               "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21]}}
t0 = time.time()
theta, z, a1, a2, rates, x = dv.table_from_json(cfg)
t1 = time.time()
rates2 = x.table(theta, z, a1, a2)
t2 = time.time()
print(json.dumps({"table": list(rates.shape), "entries": int(rates.size), "first_call_s": t1 - t0, "second_call_s": t2 - t1,
                  "kernel_ms": x.info()["last_kernel_ms"], "repeatable": bool((rates == rates2).all()), "min": float(rates.min()),
                  "max": float(rates.max())}))
rd = f1.RefTube(4, 2, 10)
t0 = time.time()
ref = f1.table(rd, rd, theta[:3], z[:2], a1[4:7], a2[5:7])
t_ref = time.time() - t0
sub = rates[:3, :2, 4:7, 5:7]
print(json.dumps({"ref_entries": int(ref.size), "ref_s": t_ref, "ref_s_per_entry": t_ref / ref.size,
                  "max_rel": float((np.abs(sub - ref) / ref).max()), "bit_equal": int((sub == ref).sum())}))
