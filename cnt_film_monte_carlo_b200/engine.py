"""numpy-facing wrapper of one libcntmc handle (include/cntmc.h).  Every method is a thin call through the C ABI."""
from __future__ import annotations

import ctypes as C
import json
from typing import Optional

import numpy as np

from . import _lib

RAND_MAX = 2147483647


class CntmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[cntmc {code}] {msg}")
        self.code = code


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Engine:
    def __init__(self, config, device: Optional[int] = None, stream: Optional[int] = None):
        """``config``: dict (whole input.json or its "exciton monte carlo" block) or JSON text."""
        self.L = _lib.load()
        text = config if isinstance(config, str) else json.dumps(config)
        h = C.c_void_p()
        rc = self.L.cntmc_create(text.encode(), C.byref(h))
        if rc != 0:
            raise CntmcError(rc, self.L.cntmc_last_error(None).decode())
        self.h = h
        if device is not None:
            self._ck(self.L.cntmc_set_device(self.h, device))
        if stream is not None:
            self._ck(self.L.cntmc_set_stream(self.h, stream))

    def close(self):
        if getattr(self, "h", None):
            self.L.cntmc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise CntmcError(rc, self.L.cntmc_last_error(self.h).decode())

    # -- inputs ---------------------------------------------------------------------------------------------------
    def load_mesh(self, directory: Optional[str] = None):
        self._ck(self.L.cntmc_load_mesh(self.h, None if directory is None else directory.encode()))

    def set_mesh(self, pos_nm: np.ndarray, orient: np.ndarray):
        _, nt, nc = pos_nm.shape
        p = np.ascontiguousarray(pos_nm.reshape(3, -1), np.float64)
        o = np.ascontiguousarray(orient.reshape(3, -1), np.float64)
        self._ck(self.L.cntmc_set_mesh(self.h, nt, nc, _p(p), _p(o)))

    def set_rate_table(self, theta, z, a1, a2, rates):
        arrs = [np.ascontiguousarray(a, np.float64) for a in (theta, z, a1, a2, rates)]
        dims = np.array(arrs[4].shape, np.int32)
        assert tuple(dims) == tuple(len(a) for a in arrs[:4])
        self._ck(self.L.cntmc_set_rate_table(self.h, _p(dims), *[_p(a) for a in arrs]))

    def rate_table(self):
        dims = np.empty(4, np.int32)
        self._ck(self.L.cntmc_get_rate_table_dims(self.h, _p(dims)))
        th, z, a1, a2 = (np.empty(int(n)) for n in dims)
        rates = np.empty(tuple(int(n) for n in dims))
        self._ck(self.L.cntmc_get_rate_table(self.h, _p(th), _p(z), _p(a1), _p(a2), _p(rates)))
        return dict(theta=th, z=z, a1=a1, a2=a2, rates=rates)

    def save_rate_table(self, directory: str):
        """scattering_struct::save (scattering_struct.h:56-94): scat_table.*.dat"""
        self._ck(self.L.cntmc_save_rate_table(self.h, directory.encode()))

    def load_rate_table(self, directory: str):
        self._ck(self.L.cntmc_load_rate_table(self.h, directory.encode()))

    def set_option(self, name: str, value: int):
        self._ck(self.L.cntmc_set_option(self.h, name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return self.L.cntmc_get_option(self.h, name.encode())

    def set_stream(self, stream: int):
        self._ck(self.L.cntmc_set_stream(self.h, stream))

    # -- Green-Kubo ------------------------------------------------------------------------------------------------
    def kubo_init(self):
        self._ck(self.L.cntmc_kubo_init(self.h))

    def kubo_create_particles(self, n: int = 0, seed: int = 1, first_global_id: int = 0):
        self._ck(self.L.cntmc_kubo_create_particles(self.h, n, seed, first_global_id))

    def kubo_create_particles_replay(self, offsets, draws, logs=None):
        off = np.ascontiguousarray(offsets, np.int64)
        dr = np.ascontiguousarray(draws, np.int32)
        lg = None if logs is None else np.ascontiguousarray(logs, np.float64)
        self._ck(self.L.cntmc_kubo_create_particles_replay(self.h, len(off) - 1, _p(off), _p(dr), _p(lg)))

    def kubo_step(self, dt: float, nsteps: int = 1, want_msd: bool = True):
        msd = np.empty((nsteps, 3)) if want_msd else None
        self._ck(self.L.cntmc_kubo_step(self.h, dt, nsteps, _p(msd)))
        return msd

    def kubo_step_dev(self, dt: float, nsteps: int, dev_ptr: int):
        """Asynchronous: sums [nsteps][4] are left at device address ``dev_ptr`` (e.g. a torch tensor's data_ptr())."""
        self._ck(self.L.cntmc_kubo_step_dev(self.h, dt, nsteps, dev_ptr))

    def kubo_step_host_state(self, dt: float, nsteps: int, state: dict, want_msd: bool = True):
        """``state`` as returned by :meth:`particles` (arrays are updated in place)."""
        msd = np.empty((nsteps, 3)) if want_msd else None
        P = len(state["site"])
        self._ck(self.L.cntmc_kubo_step_host_state(
            self.h, dt, nsteps, P, _p(state["site"]), _p(state["pos"]), _p(state["delta"]), _p(state["ff"]),
            _p(state["heading"]), _p(state["ndraw"]), _p(msd)))
        return msd

    def time(self) -> float:
        return self.L.cntmc_time(self.h)

    def kubo_max_time(self) -> float:
        return self.L.cntmc_kubo_max_time(self.h)

    def time_step(self) -> float:
        return self.L.cntmc_time_step(self.h)

    def number_of_particles(self) -> int:
        return self.L.cntmc_number_of_particles(self.h)

    def hops(self) -> int:
        return self.L.cntmc_hops(self.h)

    def reinjections(self) -> int:
        return self.L.cntmc_reinjections(self.h)

    def crossings(self) -> int:
        return self.L.cntmc_crossings(self.h)

    def probes(self) -> int:
        return self.L.cntmc_probes(self.h)

    def last_step_ms(self) -> float:
        return self.L.cntmc_last_step_ms(self.h)

    def last_step_launches(self) -> int:
        return self.L.cntmc_last_step_launches(self.h)

    def sync(self):
        self._ck(self.L.cntmc_sync(self.h))

    def kernel_ms(self) -> float:
        return self.L.cntmc_last_kernel_ms(self.h)

    def kernel_launches(self) -> int:
        return self.L.cntmc_last_kernel_launches(self.h)

    # -- contacts --------------------------------------------------------------------------------------------------
    def init(self, c1_pop: int = 1100, c2_pop: int = 0, seed: int = 1, capacity: int = 0):
        self._ck(self.L.cntmc_init(self.h, c1_pop, c2_pop, seed, capacity))

    def number_of_segments(self) -> int:
        return self.L.cntmc_number_of_segments(self.h)

    def init_replay(self, c1_pop: int, c2_pop: int, offsets, draws, logs=None):
        """monte_carlo::init with the reference's recorded draws per exciton id (ids in order of birth)"""
        off = np.ascontiguousarray(offsets, np.int64)
        dr = np.ascontiguousarray(draws, np.int32)
        lg = None if logs is None else np.ascontiguousarray(logs, np.float64)
        self._ck(self.L.cntmc_init_replay(self.h, c1_pop, c2_pop, len(off) - 1, _p(off), _p(dr), None if lg is None else _p(lg)))

    def gids(self) -> np.ndarray:
        g = np.empty(self.number_of_particles(), np.uint64)
        self._ck(self.L.cntmc_get_gids(self.h, _p(g)))
        return g

    def step(self, dt: float, nsteps: int = 1):
        n = self.number_of_segments()
        pop, cur = np.empty((nsteps, n), np.int64), np.empty((nsteps, n - 1), np.int64)
        self._ck(self.L.cntmc_step(self.h, dt, nsteps, _p(pop), _p(cur)))
        return pop, cur

    def step_dev(self, dt: float, nsteps: int, dev_ptr: int):
        self._ck(self.L.cntmc_step_dev(self.h, dt, nsteps, dev_ptr))

    def area(self):
        a = np.empty(self.number_of_segments())
        self._ck(self.L.cntmc_get_area(self.h, _p(a)))
        return a

    def scatterer_statistics(self):
        """monte_carlo::get_scatterer_statistics (monte_carlo.h:691-719): sites per slab"""
        pop = np.empty(self.number_of_segments(), np.int64)
        self._ck(self.L.cntmc_get_scatterer_statistics(self.h, _p(pop)))
        return pop

    def track_particle(self, dt: float, seed: int = 1, global_id: int = 0, max_steps: int = 1 << 20, replay_draws=None,
                       replay_logs=None):
        """monte_carlo::track_particle (monte_carlo.h:786-818): (path [n][3], reached)"""
        path = np.empty((max_steps, 3))
        n, reached = C.c_int64(), C.c_int32()
        if replay_draws is not None:
            dr = np.ascontiguousarray(replay_draws, np.int32)
            lg = None if replay_logs is None else np.ascontiguousarray(replay_logs, np.float64)
            self._ck(self.L.cntmc_track_particle(self.h, dt, 0, 0, len(dr), _p(dr), None if lg is None else _p(lg), max_steps,
                                                 _p(path), C.byref(n), C.byref(reached)))
        else:
            self._ck(self.L.cntmc_track_particle(self.h, dt, seed, global_id, 0, None, None, max_steps, _p(path), C.byref(n),
                                                 C.byref(reached)))
        return path[:n.value].copy(), bool(reached.value)

    def contact_sites(self, which: int):
        n = C.c_int64()
        self._ck(self.L.cntmc_num_contact_sites(self.h, which, C.byref(n)))
        ids = np.empty(n.value, np.int32)
        self._ck(self.L.cntmc_get_contact_sites(self.h, which, _p(ids)))
        return ids

    # -- read-back -------------------------------------------------------------------------------------------------
    def num_sites(self) -> int:
        n = C.c_int64()
        self._ck(self.L.cntmc_num_sites(self.h, C.byref(n)))
        return n.value

    def sites(self):
        N = self.num_sites()
        pos, ori = np.empty((3, N)), np.empty((3, N))
        left, right = np.empty(N, np.int32), np.empty(N, np.int32)
        rate, inv = np.empty(N), np.empty(N)
        self._ck(self.L.cntmc_get_sites(self.h, _p(pos), _p(ori), _p(left), _p(right), _p(rate), _p(inv)))
        return dict(pos=pos, orient=ori, left=left, right=right, max_rate=rate, inv_max_rate=inv)

    def domain(self):
        d = np.empty(6)
        self._ck(self.L.cntmc_get_domain(self.h, _p(d)))
        return d

    def removal_domain(self):
        d = np.empty(6)
        self._ck(self.L.cntmc_get_removal_domain(self.h, _p(d)))
        return d

    def inject(self):
        n = C.c_int64()
        self._ck(self.L.cntmc_num_inject(self.h, C.byref(n)))
        ids = np.empty(n.value, np.int32)
        self._ck(self.L.cntmc_get_inject(self.h, _p(ids)))
        return ids

    def csr(self):
        nnz = C.c_int64()
        self._ck(self.L.cntmc_csr_nnz(self.h, C.byref(nnz)))
        rp = np.empty(self.num_sites() + 1, np.int64)
        nbr, cum = np.empty(nnz.value, np.int32), np.empty(nnz.value)
        self._ck(self.L.cntmc_get_csr(self.h, _p(rp), _p(nbr), _p(cum)))
        return rp, nbr, cum

    def csr_midpoint_guards(self) -> int:
        return self.L.cntmc_csr_midpoint_guards(self.h)

    def csr_nnz(self) -> int:
        n = C.c_int64()
        self._ck(self.L.cntmc_csr_nnz(self.h, C.byref(n)))
        return n.value

    def csr_row(self, site: int, cap: int = 8192):
        """(nbr, cum) of one site's row (scatterer::find_neighbors of that site)"""
        nbr, cum, n = np.empty(cap, np.int32), np.empty(cap), C.c_int64()
        self._ck(self.L.cntmc_get_csr_row(self.h, site, cap, _p(nbr), _p(cum), C.byref(n)))
        assert n.value <= cap
        return nbr[:n.value].copy(), cum[:n.value].copy()

    def csr_build_seconds(self) -> float:
        return self.L.cntmc_csr_build_seconds(self.h)

    def particles(self):
        P = self.number_of_particles()
        site, heading, ndraw = np.empty(P, np.int32), np.empty(P, np.uint8), np.empty(P, np.uint32)
        pos, delta, ff = np.empty((3, P)), np.empty((3, P)), np.empty(P)
        self._ck(self.L.cntmc_get_particles(self.h, _p(site), _p(pos), _p(delta), _p(ff), _p(heading), _p(ndraw)))
        return dict(site=site, pos=pos, delta=delta, ff=ff, heading=heading, ndraw=ndraw)

    def trace_enable(self, cap: int):
        self._ck(self.L.cntmc_trace_enable(self.h, cap))
        self._trace_cap = cap

    def trace(self):
        P = self.number_of_particles()
        counts, sites = np.empty(P, np.int32), np.empty((P, self._trace_cap), np.int32)
        self._ck(self.L.cntmc_trace_get(self.h, _p(counts), _p(sites)))
        return counts, sites


class MultiEngine:
    """One simulation on several GPUs of one box (cntmc_multi_*, include/cntmc.h): tables replicated, excitons split by
    global id, one NCCL all-reduce of the per-step rows per call -- all inside libcntmc.so, one host thread."""

    def __init__(self, config, devices):
        self.L = _lib.load()
        text = config if isinstance(config, str) else json.dumps(config)
        devs = np.ascontiguousarray(list(devices), np.int32)
        m = C.c_void_p()
        rc = self.L.cntmc_multi_create(text.encode(), len(devs), _p(devs), C.byref(m))
        if rc != 0:
            raise CntmcError(rc, self.L.cntmc_multi_last_error(None).decode())
        self.m = m
        self.n = len(devs)
        self.n_seg = 0

    def close(self):
        if getattr(self, "m", None):
            self.L.cntmc_multi_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int):
        if rc != 0:
            raise CntmcError(rc, self.L.cntmc_multi_last_error(self.m).decode())

    def nccl_version(self) -> int:
        return self.L.cntmc_multi_nccl_version()

    def set_mesh(self, pos_nm: np.ndarray, orient: np.ndarray):
        _, nt, nc = pos_nm.shape
        p = np.ascontiguousarray(pos_nm.reshape(3, -1), np.float64)
        o = np.ascontiguousarray(orient.reshape(3, -1), np.float64)
        self._ck(self.L.cntmc_multi_set_mesh(self.m, nt, nc, _p(p), _p(o)))

    def load_mesh(self, directory: Optional[str] = None):
        self._ck(self.L.cntmc_multi_load_mesh(self.m, None if directory is None else directory.encode()))

    def set_option(self, name: str, value: int):
        self._ck(self.L.cntmc_multi_set_option(self.m, name.encode(), int(value)))

    def kubo_init(self):
        self._ck(self.L.cntmc_multi_kubo_init(self.m))

    def kubo_create_particles(self, n: int = 0, seed: int = 1):
        self._ck(self.L.cntmc_multi_kubo_create_particles(self.m, n, seed))

    def kubo_step(self, dt: float, nsteps: int) -> np.ndarray:
        msd = np.empty((nsteps, 3))
        self._ck(self.L.cntmc_multi_kubo_step(self.m, dt, nsteps, _p(msd)))
        return msd

    def particles(self):
        P = self.L.cntmc_multi_number_of_particles(self.m)
        site, heading, ndraw = np.empty(P, np.int32), np.empty(P, np.uint8), np.empty(P, np.uint32)
        pos, delta, ff = np.empty((3, P)), np.empty((3, P)), np.empty(P)
        self._ck(self.L.cntmc_multi_get_particles(self.m, _p(site), _p(pos), _p(delta), _p(ff), _p(heading), _p(ndraw)))
        return dict(site=site, pos=pos, delta=delta, ff=ff, heading=heading, ndraw=ndraw)

    def hops(self) -> int:
        return self.L.cntmc_multi_hops(self.m)

    def time(self) -> float:
        return self.L.cntmc_multi_time(self.m)

    def number_of_particles(self) -> int:
        return self.L.cntmc_multi_number_of_particles(self.m)

    def init(self, c1_pop: int, c2_pop: int, seed: int = 1):
        self._ck(self.L.cntmc_multi_init(self.m, c1_pop, c2_pop, seed))
        self.n_seg = self.L.cntmc_number_of_segments(C.c_void_p(self.L.cntmc_multi_handle(self.m, 0)))

    def step(self, dt: float, nsteps: int):
        pop = np.empty((nsteps, self.n_seg), np.int64)
        cur = np.empty((nsteps, self.n_seg - 1), np.int64)
        self._ck(self.L.cntmc_multi_step(self.m, dt, nsteps, _p(pop), _p(cur)))
        return pop, cur
