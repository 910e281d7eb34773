"""Golden vectors of the davoody rate table from the reference's own code.

Runs in the development container only: it needs oracle/_ref/libf1.so, i.e. /root/reference's exciton_transfer/cnt.cpp and
exciton_transfer/exciton_transfer.cpp compiled by oracle/Makefile (target f1) against oracle/arma_full/armadillo.  Writes
tests/golden/davoody.npz:

    tube_<n>_<m>_<cells>_<kind>      exciton dispersion energy[nk_cm, n_principal] in joules, kind in {A1, A2s, A2t}
    tube_<n>_<m>_<cells>_meta        [radius, length_in_meter, Au, ik_cm of row 0, electron-hole pairs per state]
    case_<i>_spec                    [n_d, m_d, cells_d, n_a, m_a, cells_a, temperature K, broadening meV]
    case_<i>_placements              [k, 4] = z shift, axis shift 1, axis shift 2, theta (metres, radians)
    case_<i>_rates                   [k]    = exciton_transfer::first_order of the reference, 1/s

    python tests/golden/make_golden_davoody.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import f1  # noqa: E402

PI = 3.141592  # the reference's
TUBES = [(4, 2, 10), (5, 3, 4), (8, 0, 11), (4, 2, 20), (6, 5, 3)]
grid_corners = [(1.5e-9, -10e-9, -10e-9, 0.0), (1.5e-9, 0.0, 0.0, 0.0), (1.5e-9, 0.0, 0.0, PI / 2), (1.5e-9, 0.0, 0.0, PI),
                (10e-9, 10e-9, 10e-9, PI), (2.35e-9, -2e-9, 4e-9, 9 * (PI / 180)), (5.75e-9, 10e-9, -10e-9, 171 * (PI / 180))]
CASES = [  # donor, acceptor, temperature, broadening meV, number of random placements, fixed placements
    ((4, 2, 10), (4, 2, 10), 300.0, 4.0, 6, grid_corners),  # the shipped input.json's tube: 3 K_cm, one pass of 4
    ((4, 2, 20), (4, 2, 20), 300.0, 4.0, 4, []),            # 5 K_cm: one pass of 8
    ((4, 2, 40), (4, 2, 40), 300.0, 4.0, 3, []),            # 11 K_cm: one pass of 16
    ((4, 2, 40), (4, 2, 40), 900.0, 4.0, 2, []),            # 23 K_cm: two passes of 16
    ((4, 2, 10), (4, 2, 14), 300.0, 4.0, 4, []),            # tubes of different length
    ((5, 3, 4), (5, 3, 4), 300.0, 6.5, 4, []),              # another chirality, another broadening
    ((4, 2, 6), (5, 3, 4), 300.0, 4.0, 2, []),              # no energy-matched states: all rates are zero
    ((4, 2, 10), (5, 3, 4), 300.0, 4.0, 4, []),             # two different chiralities, downhill: rates of 1e10 /s
    ((4, 2, 10), (6, 5, 3), 300.0, 4.0, 3, []),             # far apart in energy: only the Lorentzian's tail is left (1e-19 /s)
]


def main():
    out = {}
    tubes = {}

    def tube(spec):
        if spec not in tubes:
            tubes[spec] = f1.RefTube(*spec)
        return tubes[spec]

    for spec in TUBES:
        t = tube(spec)
        name = "tube_%d_%d_%d" % spec
        for kind, tag in ((f1.A1, "A1"), (f1.A2_SINGLET, "A2s"), (f1.A2_TRIPLET, "A2t")):
            e, ik0, nkc = t.exciton_energy(kind)
            out[name + "_" + tag] = e
        out[name + "_meta"] = np.array([t.radius, t.length_in_meter, t.Au, ik0, nkc], np.float64)
    rng = np.random.default_rng(20261018)
    for i, (d, a, temp, broad, n_random, fixed) in enumerate(CASES):
        td, ta = tube(d), tube(a)
        placements = list(fixed)
        for _ in range(n_random):
            placements.append((rng.uniform(1.5e-9, 10e-9), rng.uniform(-10e-9, 10e-9), rng.uniform(-10e-9, 10e-9), rng.uniform(0, PI)))
        placements = np.array(placements, np.float64)
        rates = np.array([f1.first_order_at(td, ta, temp, broad, *p) for p in placements])
        out["case_%d_spec" % i] = np.array([*d, *a, temp, broad], np.float64)
        out["case_%d_placements" % i] = placements
        out["case_%d_rates" % i] = rates
        print("case", i, d, a, temp, broad, rates[:3])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "davoody.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
