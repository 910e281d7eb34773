"""ctypes wrapper of tests/host_emul/libemul.so: the engine's host/device-shared arithmetic compiled for the CPU.

TEST INFRASTRUCTURE ONLY -- it lets `-m "not gpu"` tests exercise csrc/hop_core.h, csrc/csr_core.h and
csrc/host_setup.cpp (the code the CUDA kernels inline) against the oracle without a GPU.
"""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "host_emul", "libemul.so")
SRC = [os.path.join(HERE, "host_emul", "emul.cpp"), os.path.join(ROOT, "cnt_film_monte_carlo_b200", "csrc", "host_setup.cpp")]
DEPS = SRC + [os.path.join(ROOT, "cnt_film_monte_carlo_b200", "csrc", f) for f in ("hop_core.h", "csr_core.h", "host_setup.h", "json_min.h")]


def build():
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", LIB, *SRC])
    return LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Emul:
    def __init__(self, mc_block: dict):
        L = C.CDLL(build())
        V, I64, D, U64 = C.c_void_p, C.c_int64, C.c_double, C.c_uint64
        L.emul_create.restype = V
        L.emul_create.argtypes = [C.c_char_p]
        L.emul_error.restype = C.c_char_p
        L.emul_error.argtypes = [V]
        L.emul_destroy.argtypes = [V]
        L.emul_kubo_init.argtypes = [V, I64, I64, V, V]
        for n in ("emul_num_sites", "emul_nnz", "emul_guards", "emul_num_inject", "emul_hops", "emul_top_events"):
            getattr(L, n).restype = I64
            getattr(L, n).argtypes = [V]
        L.emul_sites.argtypes = [V] * 7
        L.emul_csr.argtypes = [V] * 4
        L.emul_domains.argtypes = [V] * 3
        L.emul_inject.argtypes = [V, V]
        L.emul_table.argtypes = [V, V]
        L.emul_create_philox.argtypes = [V, I64, U64, U64]
        L.emul_create_replay.argtypes = [V, I64, V, V, V]
        L.emul_kubo_step.argtypes = [V, D, I64, V, C.c_int]
        L.emul_particles.argtypes = [V] * 7
        L.emul_trace_counts.argtypes = [V, V]
        L.emul_trace.argtypes = [V, V]
        L.emul_select.restype = I64
        L.emul_select.argtypes = [V, I64, D]
        self.L = L
        self.h = L.emul_create(json.dumps({"exciton monte carlo": mc_block}).encode())
        err = L.emul_error(self.h).decode()
        if err:
            raise ValueError(err)
        self.P = 0

    def __del__(self):
        try:
            self.L.emul_destroy(self.h)
        except Exception:
            pass

    def kubo_init(self, pos_nm, orient):
        _, nt, nc = pos_nm.shape
        p = np.ascontiguousarray(pos_nm.reshape(3, -1))
        o = np.ascontiguousarray(orient.reshape(3, -1))
        if self.L.emul_kubo_init(self.h, nt, nc, _p(p), _p(o)) != 0:
            raise ValueError(self.L.emul_error(self.h).decode())

    def sites(self):
        N = self.L.emul_num_sites(self.h)
        pos, ori = np.empty((3, N)), np.empty((3, N))
        left, right = np.empty(N, np.int32), np.empty(N, np.int32)
        rate, inv = np.empty(N), np.empty(N)
        self.L.emul_sites(self.h, _p(pos), _p(ori), _p(left), _p(right), _p(rate), _p(inv))
        return dict(pos=pos, orient=ori, left=left, right=right, max_rate=rate, inv_max_rate=inv)

    def csr(self):
        N, nnz = self.L.emul_num_sites(self.h), self.L.emul_nnz(self.h)
        rp, nbr, cum = np.empty(N + 1, np.int64), np.empty(nnz, np.int32), np.empty(nnz)
        self.L.emul_csr(self.h, _p(rp), _p(nbr), _p(cum))
        return rp, nbr, cum

    def guards(self):
        return self.L.emul_guards(self.h)

    def domains(self):
        d, r = np.empty(6), np.empty(6)
        self.L.emul_domains(self.h, _p(d), _p(r))
        return d, r

    def inject(self):
        ids = np.empty(self.L.emul_num_inject(self.h), np.int32)
        self.L.emul_inject(self.h, _p(ids))
        return ids

    def table_rates(self, n):
        r = np.empty(n)
        self.L.emul_table(self.h, _p(r))
        return r

    def create_philox(self, P, seed, first_gid=0):
        self.P = P
        self.L.emul_create_philox(self.h, P, seed, first_gid)

    def create_replay(self, offsets, draws, logs=None):
        self.P = len(offsets) - 1
        off = np.ascontiguousarray(offsets, np.int64)
        dr = np.ascontiguousarray(draws, np.int32)
        lg = None if logs is None else np.ascontiguousarray(logs, np.float64)
        self.L.emul_create_replay(self.h, self.P, _p(off), _p(dr), _p(lg))

    def kubo_step(self, dt, nsteps, trace_cap=0):
        msd = np.empty((nsteps, 3))
        bad = self.L.emul_kubo_step(self.h, dt, nsteps, _p(msd), trace_cap)
        assert bad == 0, "emulated lane got stuck or ran out of replay draws"
        return msd

    def hops(self):
        return self.L.emul_hops(self.h)

    def top_events(self):
        """events decided by the top entries of the site record (hop_core.h after_flight_scatter)"""
        return self.L.emul_top_events(self.h)

    def set_top_entries(self, on=True):
        self.L.emul_set_top_entries.argtypes = [C.c_void_p, C.c_int]
        self.L.emul_set_top_entries(self.h, 1 if on else 0)

    def set_runs(self, on=True):
        """chain walks over memory-consecutive sites on segment times (hop_core.h fly) vs record by record"""
        self.L.emul_set_runs.argtypes = [C.c_void_p, C.c_int]
        self.L.emul_set_runs(self.h, 1 if on else 0)

    def particles(self):
        P = self.P
        site, heading, ndraw = np.empty(P, np.int32), np.empty(P, np.int32), np.empty(P, np.uint32)
        pos, delta, ff = np.empty((3, P)), np.empty((3, P)), np.empty(P)
        self.L.emul_particles(self.h, _p(site), _p(pos), _p(delta), _p(ff), _p(heading), _p(ndraw))
        return dict(site=site, pos=pos, delta=delta, ff=ff, heading=heading, ndraw=ndraw)

    def trace(self):
        counts = np.empty(self.P, np.int64)
        self.L.emul_trace_counts(self.h, _p(counts))
        off = np.zeros(self.P + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        flat = np.empty(off[-1], np.int32)
        self.L.emul_trace(self.h, _p(flat))
        return off, flat

    def select(self, cum, dice):
        cum = np.ascontiguousarray(cum, np.float64)
        return self.L.emul_select(_p(cum), len(cum), dice)


def argmin_mismatches(grid, xs):
    """(mismatches, hinted) of csr_core.h's windowed nearest-grid search against the reference's full scan"""
    L = C.CDLL(build())
    L.emul_argmin_mismatches.restype = C.c_int64
    L.emul_argmin_mismatches.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]
    g, x = np.ascontiguousarray(grid, np.float64), np.ascontiguousarray(xs, np.float64)
    hinted = C.c_int()
    bad = L.emul_argmin_mismatches(_p(g), len(g), _p(x), len(x), C.byref(hinted))
    return bad, bool(hinted.value)
