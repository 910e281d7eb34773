"""Host half of the davoody rate table (csrc/davoody_tube.h through the C ABI; no GPU needed): the tube physics against the
committed reference vectors and, where oracle/_ref/libf1.so exists, against the reference's own code live."""
import numpy as np
import pytest

import davoody_case as dc
from cnt_film_monte_carlo_b200 import davoody as dv
from cnt_film_monte_carlo_b200.engine import CntmcError, Engine
from conftest import base_mc
from oracle import f1


def test_exciton_dispersions_are_the_reference_bits():
    """cnt::calculate_exciton_dispersion: geometry, bands, v(q), Pi(q), eps(q) and the three Bethe-Salpeter problems per
    K_cm end in energies that equal the reference's bit for bit (same arithmetic order, same Jacobi eigensolver)."""
    z = dc.load()
    specs = dc.tube_specs(z)
    assert len(specs) >= 5
    for spec in specs:
        t = dv.Tube(*spec)
        name = "tube_%d_%d_%d" % spec
        meta = z[name + "_meta"]
        assert (t.radius, t.length_in_meter, t.Au) == tuple(meta[:3])
        for tag, kind in dc.KINDS.items():
            e, ik0, nkc = t.exciton_energy(kind)
            ref = z[name + "_" + tag]
            assert e.shape == ref.shape and ik0 == int(meta[3]) and nkc == int(meta[4])
            assert np.array_equal(e, ref), (spec, tag, np.abs(e - ref).max())
        # physics sanity the fixture cannot give: the singlet lies above the triplet (exchange is repulsive), energies ascend in n
        s, _, _ = t.exciton_energy(dv.A2_SINGLET)
        tr, _, _ = t.exciton_energy(dv.A2_TRIPLET)
        assert np.all(s[:, 0] >= tr[:, 0]) and np.all(np.diff(s, axis=1) >= 0)


@pytest.mark.skipif(not f1.available(), reason="oracle/_ref/libf1.so not built (no /root/reference here)")
def test_tube_against_the_reference_code_live():
    for spec in [(7, 5, 3), (9, 1, 2), (6, 0, 9)]:  # chiralities the fixture does not hold
        t, r = dv.Tube(*spec), f1.RefTube(*spec)
        assert (t.radius, t.length_in_meter, t.Au) == (r.radius, r.length_in_meter, r.Au)
        for kind in (dv.A1, dv.A2_SINGLET, dv.A2_TRIPLET):
            e, ik0, nkc = t.exciton_energy(kind)
            er, ik0r, nkcr = r.exciton_energy(kind)
            assert (ik0, nkc) == (ik0r, nkcr) and np.array_equal(e, er)


def test_eigensolver_against_lapack():
    """The one piece the davoody oracle shares with the product is the Hermitian eigensolver (the reference calls LAPACK through
    arma::eig_sym, absent here; the stand-in it is compiled against uses csrc/herm_eig.h too).  Pin it to LAPACK directly:
    numpy.linalg.eigh is zheevd, the routine behind eig_sym."""
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 8, 21, 40, 64):
        for kind in range(3):
            a = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
            a = a + a.conj().T
            if kind == 1 and n > 2:   # degenerate and clustered spectra, as K_cm = 0 blocks have
                q, _ = np.linalg.qr(a)
                d = np.repeat(rng.normal(size=(n + 1) // 2), 2)[:n]
                a = (q * d) @ q.conj().T
                a = (a + a.conj().T) / 2
            if kind == 2:             # the scale of a Bethe-Salpeter kernel in joules
                a = a * 1.6e-19
            w, v = dv.hermitian_eig(a)
            w_ref = np.linalg.eigh(a)[0]
            scale = np.abs(w_ref).max()
            assert np.all(np.diff(w) >= 0)
            assert np.abs(w - w_ref).max() <= 4e-15 * n * scale
            assert np.abs(a @ v - v * w).max() <= 1e-14 * n * scale          # A V = V diag(w)
            assert np.abs(v.conj().T @ v - np.eye(n)).max() <= 1e-14 * n     # orthonormal columns


def test_lattice_numbers_of_known_tubes():
    t = dv.Tube(4, 2, 10)  # the shipped input.json's tube (Nu=28, t=(4,-5), M=6, Q=2, 140 k x 2 mu)
    assert (t.Nu, t.M, t.Q, t.nk, t.sites) == (28, 6, 2, 140, 280)
    z = dv.Tube(8, 0, 3)   # zigzag: the unit cell is one hexagon ring pair
    assert (z.Nu, z.sites) == (16, 48)
    a = dv.Tube(5, 5, 2)   # armchair
    assert a.Nu == 10


def test_invalid_tubes_are_refused():
    for bad in [(0, 0, 10), (2, 4, 10), (4, -1, 10), (4, 2, 0)]:
        with pytest.raises(CntmcError):
            dv.Tube(*bad)


def test_table_axes_follow_linspace_and_the_reference_pi():
    mc = base_mc()
    theta, z, a1, a2 = dv.table_axes(mc)
    assert (len(theta), len(z), len(a1), len(a2)) == (21, 11, 11, 11)
    assert theta[-1] == 180 * (3.141592 / 180) and theta[0] == 0 and z[0] == 1.5e-9 and z[-1] == 10e-9
    assert a1[5] == -10e-9 + 5 * ((10e-9 - -10e-9) / 10)


def test_tubes_from_json_skips_directory_and_comment():
    cnts = {"comment": "c", "directory": "~/x", "1": {"keep old results": False, "chirality": [4, 2], "length": [3, "cnt unit cells"]}}
    tubes = dv.tubes_from_json(cnts)
    assert [(t.n, t.m, t.cells) for t in tubes] == [(4, 2, 3)]
    with pytest.raises(ValueError):
        dv.tubes_from_json({"1": {"chirality": [4, 2], "length": [3, "nm"]}})


def test_davoody_without_tubes_or_table_fails_at_init(golden_small):
    mc = dict(golden_small.mc)
    mc["rate type"] = "davoody"
    e = Engine(mc)
    e.set_mesh(golden_small.pos_nm, golden_small.orient)
    with pytest.raises(CntmcError, match="cnts"):
        e.kubo_init()
    with pytest.raises(CntmcError, match="cnt unit cells"):
        Engine({"cnts": {"1": {"chirality": [4, 2], "length": [3, "nm"]}}, "exciton monte carlo": mc})
