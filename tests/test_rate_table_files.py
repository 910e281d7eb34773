"""scat_table.*.dat: scattering_struct::save (scattering_struct.h:56-94) and the loader the reference lacks.  Host-only
entry points of the C ABI: no GPU needed."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200.engine import CntmcError, Engine
from conftest import base_mc


def random_table(rng, dims=(5, 4, 3, 6)):
    th = np.linspace(0, 180, dims[0]) * (3.141592 / 180)
    z = np.linspace(1.5e-9, 10e-9, dims[1])
    a1 = np.linspace(-10e-9, 10e-9, dims[2])
    a2 = np.linspace(-10e-9, 10e-9, dims[3])
    rates = rng.random(dims) * 10.0 ** rng.integers(8, 15, dims)
    return th, z, a1, a2, rates


def test_saved_table_loads_back_bit_for_bit(tmp_path):
    rng = np.random.default_rng(3)
    th, z, a1, a2, rates = random_table(rng)
    e = Engine(base_mc())
    e.set_rate_table(th, z, a1, a2, rates)
    e.save_rate_table(str(tmp_path))
    e2 = Engine(base_mc(**{"rate type": "davoody"}))      # a davoody run is fed from the files
    e2.load_rate_table(str(tmp_path))
    got = e2.rate_table()
    for k, v in dict(theta=th, z=z, a1=a1, a2=a2, rates=rates).items():
        assert np.array_equal(got[k], v), k


def test_files_have_the_reference_layout(tmp_path):
    """read them the way visualization/monte_carlo_results.py:261-282 does"""
    rng = np.random.default_rng(4)
    th, z, a1, a2, rates = random_table(rng, (3, 2, 4, 5))
    e = Engine(base_mc())
    e.set_rate_table(th, z, a1, a2, rates)
    e.save_rate_table(str(tmp_path))
    lines = open(tmp_path / "scat_table.rates.dat").read().splitlines()
    assert lines[0] == "sizes:" and lines[1] == "theta, z_shift, axis_shift_1, axis_shift_2" and lines[3] == ""
    dims = [int(d) for d in lines[2].split(",")]
    assert dims == [3, 2, 4, 5]
    loaded = np.loadtxt(tmp_path / "scat_table.rates.dat", skiprows=4).reshape(dims, order="C")
    assert np.array_equal(loaded, rates)                  # theta-major, axis_shift_2 fastest (scattering_struct.h:82-90)
    for name, v in (("theta", th), ("z_shift", z), ("axis_shift_1", a1), ("axis_shift_2", a2)):
        assert np.array_equal(np.loadtxt(tmp_path / ("scat_table.%s.dat" % name)), v)


def test_reference_precision_files_are_accepted(tmp_path):
    """a table written by the reference itself: default ostream precision (6 digits), Armadillo-style axis columns"""
    th, z, a1, a2 = np.array([0.0, 1.5708, 3.1416]), np.array([1.5e-9, 1e-8]), np.array([-1e-8, 1e-8]), np.array([0.0])
    rates = np.arange(12, dtype=float).reshape(3, 2, 2, 1) * 1.25e11
    for name, v in (("theta", th), ("z_shift", z), ("axis_shift_1", a1), ("axis_shift_2", a2)):
        with open(tmp_path / ("scat_table.%s.dat" % name), "w") as f:
            f.write("".join("   %.4e\n" % x for x in v))
    with open(tmp_path / "scat_table.rates.dat", "w") as f:
        f.write("sizes:\ntheta, z_shift, axis_shift_1, axis_shift_2\n3,2,2,1\n\n" + "".join("%g\n" % r for r in rates.ravel()))
    e = Engine(base_mc())
    e.load_rate_table(str(tmp_path))
    got = e.rate_table()
    assert np.array_equal(got["rates"], rates) and np.allclose(got["theta"], th) and got["rates"].shape == (3, 2, 2, 1)


def test_loader_errors(tmp_path):
    e = Engine(base_mc())
    with pytest.raises(CntmcError):
        e.load_rate_table(str(tmp_path))                  # no files
    with pytest.raises(CntmcError):
        e.save_rate_table(str(tmp_path))                  # no table yet
    rng = np.random.default_rng(5)
    e.set_rate_table(*random_table(rng, (2, 2, 2, 2)))
    e.save_rate_table(str(tmp_path))
    with open(tmp_path / "scat_table.theta.dat", "a") as f:
        f.write("3.0\n")
    with pytest.raises(CntmcError, match="sizes do not match"):
        e.load_rate_table(str(tmp_path))
