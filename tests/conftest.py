import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["small_forster", "wong_trimmed"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


class Golden:
    """One committed fixture: inputs + what the reference's own code computed from them (tests/golden/make_golden.py)."""

    def __init__(self, name):
        self.name = name
        with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
            self.mc = json.load(f)["exciton monte carlo"]
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.pos_nm, self.orient = self.z["pos_nm"], self.z["orient"]
        self.P = int(self.mc["number of particles for kubo simulation"])
        self.dt = float(self.z["dt"])
        self.nsteps = int(self.z["nsteps"])
        self.seed = int(self.z["seed"])

    def group(self, prefix):
        return {k[len(prefix):]: self.z[k] for k in self.z.files if k.startswith(prefix)}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)


@pytest.fixture
def golden_small():
    return Golden("small_forster")


def wide_trim(mc):
    mc = dict(mc)
    mc["trim limits"] = {"xlim": [-1e-5, 1e-5], "ylim": [-1e-5, 1e-5], "zlim": [-1e-5, 1e-5]}
    return mc


def base_mc(**over):
    """The reference's input.json "exciton monte carlo" block with the closed-form table (input.json:40-68)."""
    mc = {
        "rate type": "forster",
        "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
        "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21],
        "temperature [kelvin]": 300, "max hopping radius [m]": 20e-9, "number of segments": 10,
        "trim limits": {"xlim": [-1e-5, 1e-5], "ylim": [-1e-5, 1e-5], "zlim": [-1e-5, 1e-5]},
        "exciton velocity [m/s]": 2e5, "monte carlo time step": 1e-13,
        "number of sections for injection region": 5,
        "maximum time for kubo simulation [seconds]": 1e-9,
        "number of particles for kubo simulation": 2000,
    }
    mc.update(over)
    return mc
