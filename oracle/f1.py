"""TEST INFRASTRUCTURE ONLY: ctypes face of oracle/_ref/libf1.so -- the reference's own cnt.cpp and exciton_transfer.cpp
(compiled by oracle/Makefile from /root/reference against oracle/arma_full/armadillo) behind oracle/f1_driver.cpp.

The reference prints progress to stdout and writes its result files below the directory it is given and below
$HOME/research; both are pointed at a scratch directory, and stdout is silenced around the calls.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libf1.so")
A1, A2_SINGLET, A2_TRIPLET = 0, 1, 2

_lib = None
_scratch = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib, _scratch
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.f1_cnt_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p]
        L.f1_first_order.argtypes = [C.c_int, C.c_int] + [C.c_double] * 4
        L.f1_first_order.restype = C.c_double
        L.f1_first_order_at.argtypes = [C.c_int, C.c_int] + [C.c_double] * 6
        L.f1_first_order_at.restype = C.c_double
        L.f1_table.argtypes = [C.c_int, C.c_int] + [C.c_int, C.c_void_p] * 4 + [C.c_void_p]
        for name in ("f1_radius", "f1_length_in_meter", "f1_Au"):
            getattr(L, name).restype = C.c_double
        L.f1_last_error.restype = C.c_char_p
        _lib = L
        _scratch = tempfile.mkdtemp(prefix="cntmc_f1_")
    return _lib


@contextlib.contextmanager
def _quiet():
    """The reference narrates on std::cout; send fd 1 to /dev/null while it runs."""
    import sys

    sys.stdout.flush()
    saved = os.dup(1)
    null = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(null, 1)
        yield
    finally:
        os.dup2(saved, 1)
        os.close(null)
        os.close(saved)


class RefTube:
    """cnt(json, dir) + calculate_exciton_dispersion() of the reference."""

    def __init__(self, n: int, m: int, length_cells: int):
        L = _load()
        home = os.environ.get("HOME")
        with _quiet():
            self.id = L.f1_cnt_create(n, m, length_cells, _scratch.encode())
        if home is not None:
            os.environ["HOME"] = home  # the driver redirects $HOME inside the process image only; keep Python's view
        if self.id < 0:
            raise RuntimeError(L.f1_last_error().decode())
        self.radius = L.f1_radius(self.id)
        self.length_in_meter = L.f1_length_in_meter(self.id)
        self.Au = L.f1_Au(self.id)

    def exciton_energy(self, which: int = A2_SINGLET):
        L = _load()
        dims = (C.c_int * 5)()
        L.f1_exciton_dims(self.id, which, dims)
        e = np.zeros((dims[0], dims[1]), np.float64)
        L.f1_exciton_energy(self.id, which, e.ctypes.data_as(C.c_void_p))
        return e, int(dims[3]), int(dims[2])


def first_order(donor: RefTube, acceptor: RefTube, z_shift: float, axis_shift_1: float, axis_shift_2: float, theta: float) -> float:
    L = _load()
    with _quiet():
        r = L.f1_first_order(donor.id, acceptor.id, z_shift, axis_shift_1, axis_shift_2, theta)
    return float(r)


def first_order_at(donor: RefTube, acceptor: RefTube, temperature: float, broadening_mev: float, z_shift: float, axis_shift_1: float,
                   axis_shift_2: float, theta: float) -> float:
    """first_order at a temperature [K] and broadening [meV]: the two-tube constructor fixes 300 K and 4 meV, everything else
    goes through the reference's JSON constructor (which cannot tell two tubes of one chirality apart)."""
    if temperature == 300.0 and broadening_mev == 4.0:
        return first_order(donor, acceptor, z_shift, axis_shift_1, axis_shift_2, theta)
    L = _load()
    with _quiet():
        r = L.f1_first_order_at(donor.id, acceptor.id, temperature, broadening_mev, z_shift, axis_shift_1, axis_shift_2, theta)
    if r < 0:
        raise RuntimeError(L.f1_last_error().decode())
    return float(r)


def table(donor: RefTube, acceptor: RefTube, theta, z_shift, axis_shift_1, axis_shift_2) -> np.ndarray:
    """rates[theta, z, a1, a2]: the loop nest of monte_carlo.cpp:114-137 (single-threaded here)."""
    L = _load()
    axes = [np.ascontiguousarray(a, np.float64) for a in (theta, z_shift, axis_shift_1, axis_shift_2)]
    out = np.zeros(tuple(len(a) for a in axes), np.float64)
    args = []
    for a in axes:
        args += [len(a), a.ctypes.data_as(C.c_void_p)]
    with _quiet():
        rc = L.f1_table(donor.id, acceptor.id, *args, out.ctypes.data_as(C.c_void_p))
    if rc != 0:
        raise RuntimeError(L.f1_last_error().decode())
    return out
