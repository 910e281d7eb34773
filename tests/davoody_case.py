"""Shared by the davoody tests: the committed fixture tests/golden/davoody.npz (made by tests/golden/make_golden_davoody.py
from the reference's own cnt.cpp / exciton_transfer.cpp)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "davoody.npz")
KINDS = {"A1": 0, "A2s": 1, "A2t": 2}


def load():
    return np.load(GOLDEN)


def tube_specs(z):
    return sorted({tuple(int(v) for v in k.split("_")[1:4]) for k in z.files if k.startswith("tube_") and k.endswith("_meta")})


def cases(z):
    out = []
    i = 0
    while "case_%d_spec" % i in z.files:
        s = z["case_%d_spec" % i]
        out.append(dict(donor=tuple(int(v) for v in s[0:3]), acceptor=tuple(int(v) for v in s[3:6]), temperature=float(s[6]),
                        broadening_mev=float(s[7]), placements=z["case_%d_placements" % i], rates=z["case_%d_rates" % i]))
        i += 1
    return out
