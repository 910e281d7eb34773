"""ctypes view of oracle/_ref/libt0.so -- the reference's own hot-path code (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module.  The library
holds one global simulation object (the reference keeps raw pointers into its own vectors), so use one
:class:`T0` at a time per process.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libt0.so")
REF_BINARY = os.path.join(_HERE, "_ref", "cnt_mc_ref")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class T0:
    def __init__(self):
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle t0` in the dev container")
        L = C.CDLL(LIB_PATH)
        L.t0_last_error.restype = C.c_char_p
        for name in ("t0_num_sites", "t0_num_inject", "t0_num_particles", "t0_total_draws"):
            getattr(L, name).restype = C.c_int64
        for name in ("t0_time", "t0_max_time", "t0_cutoff"):
            getattr(L, name).restype = C.c_double
        L.t0_get_rate.restype = C.c_double
        L.t0_get_rate.argtypes = [C.c_double] * 4
        L.t0_row.restype = C.c_int64
        L.t0_row.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
        L.t0_open.argtypes = [C.c_char_p, C.c_uint]
        L.t0_open_contacts.argtypes = [C.c_char_p, C.c_uint]
        L.t0_kubo_step_verbatim.argtypes = [C.c_double, C.c_int64, C.c_void_p, C.c_int]
        L.t0_kubo_step_logged.argtypes = [C.c_double, C.c_int64, C.c_void_p]
        L.t0_kubo_step_omp.argtypes = [C.c_double, C.c_int64]
        L.t0_kubo_step_omp.restype = C.c_int64
        L.t0_kubo_create_particles_logged.argtypes = [C.c_int64]
        L.t0_draw_counts.argtypes = [C.c_void_p, C.c_int64]
        L.t0_draws.argtypes = [C.c_void_p, C.c_int64]
        L.t0_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.t0_degrees.argtypes = [C.c_void_p]
        L.t0_num_contact_sites.restype = C.c_int64
        L.t0_contact_iteration.argtypes = [C.c_double]
        L.t0_set_contact_pops.argtypes = [C.c_uint, C.c_uint]
        self.L = L

    # -- lifecycle ----------------------------------------------------------------------------------------------
    def open(self, json_path: str, seed: int = 100) -> None:
        if self.L.t0_open(json_path.encode(), seed) != 0:
            raise RuntimeError(self.L.t0_last_error().decode())

    def open_contacts(self, json_path: str, seed: int = 100) -> None:
        if self.L.t0_open_contacts(json_path.encode(), seed) != 0:
            raise RuntimeError(self.L.t0_last_error().decode())

    def close(self) -> None:
        self.L.t0_close()

    def srand(self, seed: int) -> None:
        self.L.t0_srand(seed)

    def set_threads(self, n: int) -> None:
        self.L.t0_set_threads(n)

    # -- set-up state -------------------------------------------------------------------------------------------
    def sites(self):
        N = self.L.t0_num_sites()
        pos, ori = np.empty((3, N)), np.empty((3, N))
        left, right = np.empty(N, np.int32), np.empty(N, np.int32)
        rate, inv = np.empty(N), np.empty(N)
        self.L.t0_sites(_p(pos), _p(ori), _p(left), _p(right), _p(rate), _p(inv))
        return dict(pos=pos, orient=ori, left=left, right=right, max_rate=rate, inv_max_rate=inv)

    def table(self):
        dims = np.empty(4, np.int32)
        self.L.t0_table_dims(_p(dims))
        th, z, a1, a2 = (np.empty(int(n)) for n in dims)
        rates = np.empty(tuple(int(n) for n in dims))
        self.L.t0_table(_p(th), _p(z), _p(a1), _p(a2), _p(rates))
        return dict(theta=th, z=z, a1=a1, a2=a2, rates=rates)

    def get_rate(self, theta, z, a1, a2) -> float:
        return self.L.t0_get_rate(theta, z, a1, a2)

    def domain(self):
        d = np.empty(6)
        self.L.t0_domain(_p(d))
        return d

    def removal_domain(self):
        d = np.empty(6)
        self.L.t0_removal_domain(_p(d))
        return d

    def inject(self):
        ids = np.empty(self.L.t0_num_inject(), np.int32)
        self.L.t0_inject(_p(ids))
        return ids

    def row(self, i: int, cap: int = 4096):
        ids, cum = np.empty(cap, np.int32), np.empty(cap)
        d = self.L.t0_row(i, _p(ids), _p(cum), cap)
        assert d <= cap
        return ids[:d].copy(), cum[:d].copy()

    def csr(self, threads: int = 0):
        if threads:
            self.set_threads(threads)
        N = self.L.t0_num_sites()
        deg = np.empty(N, np.int32)
        self.L.t0_degrees(_p(deg))
        row_ptr = np.zeros(N + 1, np.int64)
        np.cumsum(deg, out=row_ptr[1:])
        ids, cum = np.empty(row_ptr[-1], np.int32), np.empty(row_ptr[-1])
        self.L.t0_csr(_p(row_ptr), _p(ids), _p(cum))
        if threads:
            self.set_threads(1)
        return row_ptr, ids, cum

    # -- excitons -----------------------------------------------------------------------------------------------
    def log_draws(self, on: bool = True) -> None:
        self.L.t0_log_draws(1 if on else 0)

    def create_particles_verbatim(self) -> None:
        self.L.t0_kubo_create_particles_verbatim()

    def create_particles_logged(self, P: int) -> None:
        self.L.t0_kubo_create_particles_logged(P)

    def particles(self):
        P = self.L.t0_num_particles()
        site, heading = np.empty(P, np.int32), np.empty(P, np.int32)
        pos, old, delta, ff = np.empty((3, P)), np.empty((3, P)), np.empty((3, P)), np.empty(P)
        self.L.t0_particles(_p(site), _p(pos), _p(old), _p(delta), _p(ff), _p(heading))
        return dict(site=site, pos=pos, old_pos=old, delta=delta, ff=ff, heading=heading)

    def kubo_step_verbatim(self, dt: float, nsteps: int, write_file: bool = False):
        msd = np.empty((nsteps, 3))
        self.L.t0_kubo_step_verbatim(dt, nsteps, _p(msd), 1 if write_file else 0)
        return msd

    def kubo_step_logged(self, dt: float, nsteps: int):
        msd = np.empty((nsteps, 3))
        self.L.t0_kubo_step_logged(dt, nsteps, _p(msd))
        return msd

    def kubo_step_omp(self, dt: float, nsteps: int) -> int:
        """The reference's OpenMP particle loop; returns the number of re-injections (for exact hop counting)."""
        return self.L.t0_kubo_step_omp(dt, nsteps)

    def draws(self, P: int):
        """Per-exciton draw log in CSR form: ``(offsets[P+1], flat int32 draws)``."""
        counts = np.empty(P, np.int64)
        self.L.t0_draw_counts(_p(counts), P)
        off = np.zeros(P + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        flat = np.empty(off[-1], np.int32)
        self.L.t0_draws(_p(flat), P)
        return off, flat

    def total_draws(self) -> int:
        return self.L.t0_total_draws()

    def time(self) -> float:
        return self.L.t0_time()

    # -- contact mode -------------------------------------------------------------------------------------------
    def contact_sites(self, which: int):
        ids = np.empty(self.L.t0_num_contact_sites(which), np.int32)
        self.L.t0_contact_sites(which, _p(ids))
        return ids

    def area(self, n_seg: int):
        a = np.empty(n_seg)
        self.L.t0_area(_p(a))
        return a

    def contact_iteration(self, dt: float) -> None:
        self.L.t0_contact_iteration(dt)

    def open_contacts_logged(self, json_path: str, seed: int = 100) -> int:
        """monte_carlo::init with the creation draws attributed per exciton; returns the initial population."""
        self.L.t0_open_contacts_logged.restype = C.c_int64
        self.L.t0_contacts_attribute_creation.restype = C.c_int64
        if self.L.t0_open_contacts_logged(json_path.encode(), seed) != 0:
            raise RuntimeError(self.L.t0_last_error().decode())
        p0 = self.L.t0_contacts_attribute_creation()
        if p0 < 0:
            raise RuntimeError(self.L.t0_last_error().decode())
        return p0

    def contact_iteration_logged(self, dt: float) -> None:
        """step / save_metrics / repopulate_contacts with every draw attributed to the exciton (id = order of birth)"""
        self.L.t0_contact_iteration_logged.argtypes = [C.c_double]
        self.L.t0_contact_iteration_logged(dt)

    def next_id(self) -> int:
        self.L.t0_next_id.restype = C.c_int64
        return self.L.t0_next_id()

    def particle_ids(self):
        ids = np.empty(self.L.t0_num_particles(), np.int64)
        self.L.t0_particle_ids(_p(ids))
        return ids

    def track_particle(self, dt: float, file_no: int, log_slot: int = 0) -> None:
        """monte_carlo::track_particle (monte_carlo.h:786-818); writes particle_path.<file_no>.dat in the output directory."""
        self.L.t0_track_particle.argtypes = [C.c_double, C.c_int, C.c_int64]
        self.L.t0_track_particle.restype = None
        self.L.t0_track_particle(dt, file_no, log_slot)
