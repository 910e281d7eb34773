"""The "davoody" rate table (include/cntmc.h, section "davoody rate table"): thin wrappers of the C ABI.

``Tube`` is the reference's ``cnt`` after ``calculate_exciton_dispersion`` (exciton_transfer/cnt.cpp:1056-1081, host only),
``Transfer`` its ``exciton_transfer(cnt1, cnt2)`` (exciton_transfer.h:37-48) whose ``first_order`` runs on the GPU for any
number of placements at once, ``table_from_json`` the ``"rate type": "davoody"`` branch of
``monte_carlo::create_scattering_table`` (monte_carlo.cpp:24-49, 64-153).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import CntmcError, _p

EV = 1.6 * 10 ** -19.0  # the reference's own electron volt, formed like constants.h:13 forms it
PI = 3.141592  # and its own pi (constants.h:10): degrees -> radians in the table axes use it
A1, A2_SINGLET, A2_TRIPLET = 0, 1, 2


def _fail(L, code=-1):
    raise CntmcError(code, L.cntmc_davoody_last_error().decode())


def fp64_peak(device: int = 0) -> float:
    """Measured FP64 fused multiply-add peak of a device in TFLOP/s: the denominator of the table kernel's roofline."""
    L = _lib.load()
    out = np.zeros(1, np.float64)
    if L.cntmc_fp64_peak(device, _p(out)) != 0:
        _fail(L)
    return float(out[0])


def hermitian_eig(a: np.ndarray):
    """(w ascending, V with eigenvectors in its columns) of a complex Hermitian matrix: the solver that stands in for
    arma::eig_sym (cnt.cpp:950-962) in the tube physics, csrc/herm_eig.h."""
    L = _lib.load()
    a = np.ascontiguousarray(a, np.complex128)
    n = a.shape[0]
    assert a.shape == (n, n)
    w = np.zeros(n, np.float64)
    v = np.zeros((n, n), np.complex128)
    if L.cntmc_hermitian_eig(n, _p(a.view(np.float64)), _p(w), _p(v.view(np.float64))) != 0:
        _fail(L)
    return w, v


class Tube:
    def __init__(self, n: int, m: int, length_cells: int):
        self.L = _lib.load()
        self.h = self.L.cntmc_tube_create(n, m, length_cells)
        if not self.h:
            _fail(self.L)
        ints = np.zeros(8, np.int32)
        reals = np.zeros(4, np.float64)
        if self.L.cntmc_tube_info(self.h, _p(ints), _p(reals)) != 0:
            _fail(self.L)
        self.n, self.m, self.cells, self.Nu, self.M, self.Q, self.nk, self.sites = (int(v) for v in ints)
        self.radius, self.length_in_meter, self.Au, self.build_seconds = (float(v) for v in reals)

    def close(self):
        if getattr(self, "h", None):
            self.L.cntmc_tube_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def exciton_energy(self, which: int = A2_SINGLET):
        """(energy[nk_cm, n_principal] in joules, ik_cm of row 0, number of electron-hole pairs per state)"""
        dims = np.zeros(4, np.int32)
        if self.L.cntmc_tube_exciton_dims(self.h, which, _p(dims)) != 0:
            _fail(self.L)
        e = np.zeros((int(dims[0]), int(dims[1])), np.float64)
        if self.L.cntmc_tube_exciton_energy(self.h, which, _p(e)) != 0:
            _fail(self.L)
        return e, int(dims[3]), int(dims[2])


class Transfer:
    def __init__(self, donor: Tube, acceptor: Tube, temperature: float = 300.0, broadening: float = 4.0e-3 * EV, device: int = -1):
        self.L = _lib.load()
        self.donor, self.acceptor = donor, acceptor  # the handle refers to both tubes: keep them alive
        self.h = self.L.cntmc_transfer_create(donor.h, acceptor.h, temperature, broadening, device)
        if not self.h:
            _fail(self.L, -2)

    def close(self):
        if getattr(self, "h", None):
            self.L.cntmc_transfer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self) -> dict:
        ints = np.zeros(8, np.int32)
        reals = np.zeros(4, np.float64)
        if self.L.cntmc_transfer_info(self.h, _p(ints), _p(reals)) != 0:
            _fail(self.L)
        keys = ("donor_states", "acceptor_states", "pairs", "donor_kcm", "acceptor_kcm", "kcm_per_pass", "threads", "smem_bytes")
        out = {k: int(v) for k, v in zip(keys, ints)}
        out.update(temperature=float(reals[0]), broadening=float(reals[1]), last_kernel_ms=float(reals[2]), launches=int(reals[3]))
        return out

    def pair_factors(self):
        n = self.info()["pairs"]
        q = np.zeros((n, 2), np.float64)
        b = np.zeros(n, np.float64)
        l = np.zeros(n, np.float64)
        if self.L.cntmc_transfer_pair_factors(self.h, _p(q), _p(b), _p(l)) != 0:
            _fail(self.L)
        return q[:, 0] + 1j * q[:, 1], b, l

    def first_order(self, z_shift, axis_shift_1, axis_shift_2, theta) -> np.ndarray:
        """exciton_transfer::first_order for arrays of placements (broadcast against each other); rates in 1/s."""
        z, a1, a2, th = (np.ascontiguousarray(a, np.float64) for a in np.broadcast_arrays(z_shift, axis_shift_1, axis_shift_2, theta))
        rate = np.zeros(z.shape, np.float64)
        rc = self.L.cntmc_transfer_first_order(self.h, z.size, _p(z), _p(a1), _p(a2), _p(th), _p(rate))
        if rc != 0:
            _fail(self.L, rc)
        return rate

    def table(self, theta, z_shift, axis_shift_1, axis_shift_2) -> np.ndarray:
        """rates[theta, z, a1, a2] over the four axes (theta in radians), the loop nest of monte_carlo.cpp:114-137."""
        axes = [np.ascontiguousarray(a, np.float64) for a in (theta, z_shift, axis_shift_1, axis_shift_2)]
        dims = np.array([len(a) for a in axes], np.int32)
        rates = np.zeros(tuple(int(d) for d in dims), np.float64)
        rc = self.L.cntmc_transfer_table(self.h, _p(dims), *[_p(a) for a in axes], _p(rates))
        if rc != 0:
            _fail(self.L, rc)
        return rates

    def install(self, engine, theta, z_shift, axis_shift_1, axis_shift_2):
        """Build the table and make it ``engine``'s scattering table (before kubo_init / init)."""
        axes = [np.ascontiguousarray(a, np.float64) for a in (theta, z_shift, axis_shift_1, axis_shift_2)]
        dims = np.array([len(a) for a in axes], np.int32)
        rc = self.L.cntmc_create_davoody_table(engine.h, self.h, _p(dims), *[_p(a) for a in axes])
        if rc != 0:
            _fail(self.L, rc)


def linspace(start: float, stop: float, num: int) -> np.ndarray:
    """arma::linspace as the reference's Armadillo evaluates it: start + i*delta, the last point exactly ``stop``."""
    num = int(num)
    if num < 2:
        return np.array([float(stop)])
    delta = (float(stop) - float(start)) / float(num - 1)
    out = float(start) + np.arange(num, dtype=np.float64) * delta
    out[-1] = float(stop)
    return out


def table_axes(mc_block: dict):
    """The four axes of create_davoody_scatt_table (monte_carlo.cpp:65-75) from an "exciton monte carlo" block."""
    theta = linspace(*mc_block["theta [degrees]"]) * (PI / 180)
    return theta, linspace(*mc_block["zshift [m]"]), linspace(*mc_block["axis shift 1 [m]"]), linspace(*mc_block["axis shift 2 [m]"])


def tubes_from_json(cnts_block: dict):
    """One Tube per entry of the JSON's "cnts" (monte_carlo.cpp:33-47; "directory" and "comment" are not tubes)."""
    tubes = []
    for key, spec in cnts_block.items():
        if key in ("directory", "comment"):
            continue
        length, units = spec["length"]
        if units != "cnt unit cells":
            raise ValueError('units other than "cnt unit cells" is not implemented yet!!!')  # cnt.h:185-188
        n, m = spec["chirality"]
        tubes.append(Tube(int(n), int(m), int(length)))
    return tubes


def table_from_json(config: dict, device: int = -1):
    """``config`` = a whole input.json with "rate type": "davoody".  Returns (theta, z, a1, a2, rates, transfer); the
    reference always transfers from the first tube to itself (monte_carlo.cpp:48)."""
    tubes = tubes_from_json(config["cnts"])
    if not tubes:
        raise ValueError('"cnts" names no tube')
    x = Transfer(tubes[0], tubes[0], device=device)
    theta, z, a1, a2 = table_axes(config["exciton monte carlo"])
    return theta, z, a1, a2, x.table(theta, z, a1, a2), x
