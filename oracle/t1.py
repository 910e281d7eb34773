"""ctypes view of oracle/_ref/liboracle_t1.so -- our plain-C restatement of the hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this module; the product
package never does.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "liboracle_t1.so")

GLIBC, PHILOX, REPLAY = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (seconds).  Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle_t1.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "t1"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _lib():
    build()
    L = C.CDLL(LIB_PATH)
    V, I64, I32, D, U64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double, C.c_uint64
    sig = {
        "t1_linspace": (None, [D, D, I64, V]),
        "t1_philox4x32_10": (None, [V, V, V]),
        "t1_philox2x32_10": (None, [V, C.c_uint32, V]),
        "t1_philox_draw": (I32, [U64, U64, U64]),
        "t1_forster_table": (None, [D, V, V, V, V, V, V]),
        "t1_select": (I64, [V, I64, D]),
        "t1_log_ratios": (None, [V, I64, V]),
        "t1_create": (V, []),
        "t1_destroy": (None, [V]),
        "t1_set_table": (None, [V, V, V, V, V, V, V]),
        "t1_get_rate": (D, [V, D, D, D, D]),
        "t1_set_mesh": (None, [V, I64, I64, V, V]),
        "t1_trim": (None, [V, V, V, V]),
        "t1_find_domain": (None, [V]),
        "t1_build_buckets": (None, [V, D]),
        "t1_set_velocity": (None, [V, D]),
        "t1_set_max_rate": (None, [V]),
        "t1_set_lazy_rates": (None, [V, C.c_int]),
        "t1_injection": (None, [V, I32]),
        "t1_num_sites": (I64, [V]),
        "t1_sites": (None, [V, V, V, V, V, V, V]),
        "t1_domain": (None, [V, V]),
        "t1_removal_domain": (None, [V, V]),
        "t1_num_inject": (I64, [V]),
        "t1_inject": (None, [V, V]),
        "t1_bucket_dims": (None, [V, V]),
        "t1_row": (I64, [V, I64, V, V, I64]),
        "t1_degrees": (None, [V, V]),
        "t1_csr": (None, [V, V, V, V]),
        "t1_set_memo": (None, [V, C.c_int]),
        "t1_draws_glibc": (None, [V]),
        "t1_draws_philox": (None, [V, U64]),
        "t1_draws_replay": (None, [V, I64, V, V]),
        "t1_log_draws": (None, [V, C.c_int]),
        "t1_trace_sites": (None, [V, C.c_int]),
        "t1_kubo_create_particles": (None, [V, I64, U64]),
        "t1_kubo_step": (None, [V, D, I64, V]),
        "t1_num_particles": (I64, [V]),
        "t1_particles": (None, [V, V, V, V, V, V, V]),
        "t1_time": (D, [V]),
        "t1_hops": (I64, [V]),
        "t1_reinjections": (I64, [V]),
        "t1_event_counts": (None, [V, V]),
        "t1_draw_counts": (None, [V, V]),
        "t1_logged_draws": (None, [V, V]),
        "t1_trace_counts": (None, [V, V]),
        "t1_traced_sites": (None, [V, V]),
        "t1_replay_exhausted": (C.c_int, [V]),
        "t1_contacts_init": (None, [V, I32, I64, I64]),
        "t1_area": (None, [V, V]),
        "t1_num_contact_sites": (I64, [V, C.c_int]),
        "t1_contact_sites": (None, [V, C.c_int, V]),
        "t1_contact_iteration": (None, [V, D, V, V]),
        "t1_track_particle": (I64, [V, D, C.c_uint64, I64, V, V]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    return L


_L = None


def lib():
    global _L
    if _L is None:
        _L = _lib()
    return _L


# ---- stand-alone helpers --------------------------------------------------------------------------------------
def linspace(a: float, b: float, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().t1_linspace(a, b, n, _p(out))
    return out


def philox4x32_10(ctr, key) -> np.ndarray:
    c, k, o = np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), np.empty(4, np.uint32)
    lib().t1_philox4x32_10(_p(c), _p(k), _p(o))
    return o


def philox2x32_10(ctr, key: int) -> np.ndarray:
    c, o = np.asarray(ctr, np.uint32), np.empty(2, np.uint32)
    lib().t1_philox2x32_10(_p(c), key, _p(o))
    return o


def philox_draw(seed: int, exciton: int, k: int) -> int:
    return lib().t1_philox_draw(seed, exciton, k)


def select(cum: np.ndarray, dice: float) -> int:
    cum = np.ascontiguousarray(cum, np.float64)
    return lib().t1_select(_p(cum), len(cum), dice)


def log_ratios(draws: np.ndarray) -> np.ndarray:
    """log(r/RAND_MAX) for each draw, computed by the host libm (what the reference's ff_time evaluates)."""
    d = np.ascontiguousarray(draws, np.int32)
    out = np.empty(len(d))
    lib().t1_log_ratios(_p(d), len(d), _p(out))
    return out


REF_PI = 3.141592  # helper/constants.h:10 (truncated on purpose: it defines the theta grid)


def table_grids(mc: dict):
    """The four grids of monte_carlo.cpp:157-167 from the "exciton monte carlo" JSON block."""
    zs, a1, a2, th = (mc[k] for k in ("zshift [m]", "axis shift 1 [m]", "axis shift 2 [m]", "theta [degrees]"))
    theta = linspace(th[0], th[1], int(th[2])) * (REF_PI / 180)
    return theta, linspace(zs[0], zs[1], int(zs[2])), linspace(a1[0], a1[1], int(a1[2])), linspace(a2[0], a2[1], int(a2[2]))


def forster_table(mc: dict):
    """monte_carlo.cpp:24-61 for "forster" / "wong"; returns dict(theta,z,a1,a2,rates)."""
    gamma0 = {"forster": 1.0e15, "wong": 1.0e13}[mc["rate type"]]
    theta, z, a1, a2 = table_grids(mc)
    dims = np.array([len(theta), len(z), len(a1), len(a2)], np.int32)
    rates = np.empty(tuple(int(d) for d in dims))
    lib().t1_forster_table(gamma0, _p(dims), _p(theta), _p(z), _p(a1), _p(a2), _p(rates))
    return dict(theta=theta, z=z, a1=a1, a2=a2, rates=rates)


class T1:
    """One oracle simulation.  ``kubo_init`` mirrors monte_carlo::kubo_init (monte_carlo.cpp:254-305)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.t1_create()
        self._keep = []

    def __del__(self):
        try:
            self.L.t1_destroy(self.h)
        except Exception:
            pass

    # -- set-up -------------------------------------------------------------------------------------------------
    def set_table(self, t: dict) -> None:
        dims = np.array(t["rates"].shape, np.int32)
        arrs = [np.ascontiguousarray(t[k], np.float64) for k in ("theta", "z", "a1", "a2", "rates")]
        self.L.t1_set_table(self.h, _p(dims), *[_p(a) for a in arrs])

    def kubo_init(self, mc: dict, pos_nm: np.ndarray, orient: np.ndarray, table: dict | None = None) -> None:
        self.set_table(table if table is not None else forster_table(mc))
        self.setup_sites(mc, pos_nm, orient)
        self.L.t1_injection(self.h, int(mc["number of sections for injection region"]))

    def setup_sites(self, mc: dict, pos_nm: np.ndarray, orient: np.ndarray) -> None:
        _, nt, ncol = pos_nm.shape
        p = np.ascontiguousarray(pos_nm.reshape(3, -1), np.float64)
        o = np.ascontiguousarray(orient.reshape(3, -1), np.float64)
        self.L.t1_set_mesh(self.h, nt, ncol, _p(p), _p(o))
        lims = [np.array(mc["trim limits"][k], np.float64) for k in ("xlim", "ylim", "zlim")]
        self.L.t1_trim(self.h, *[_p(a) for a in lims])
        self.L.t1_find_domain(self.h)
        self.L.t1_set_velocity(self.h, float(mc["exciton velocity [m/s]"]))
        self.L.t1_build_buckets(self.h, float(mc["max hopping radius [m]"]))
        self.L.t1_set_max_rate(self.h)

    def contacts_init(self, mc: dict, pos_nm, orient, table=None, c1_pop=1100, c2_pop=0) -> None:
        self.set_table(table if table is not None else forster_table(mc))
        self.setup_sites(mc, pos_nm, orient)
        self.n_seg = int(mc["number of segments"])
        self.L.t1_contacts_init(self.h, self.n_seg, c1_pop, c2_pop)

    def set_lazy_rates(self, on: bool = True) -> None:
        """Call before kubo_init: Gamma_i of a site is computed when first needed (the same value), for 5e6-site films."""
        self.L.t1_set_lazy_rates(self.h, 1 if on else 0)

    def set_memo(self, on: bool = True) -> None:
        self.L.t1_set_memo(self.h, 1 if on else 0)

    def get_rate(self, theta, z, a1, a2) -> float:
        return self.L.t1_get_rate(self.h, theta, z, a1, a2)

    def num_sites(self) -> int:
        return self.L.t1_num_sites(self.h)

    def sites(self):
        N = self.num_sites()
        pos, ori = np.empty((3, N)), np.empty((3, N))
        left, right = np.empty(N, np.int32), np.empty(N, np.int32)
        rate, inv = np.empty(N), np.empty(N)
        self.L.t1_sites(self.h, _p(pos), _p(ori), _p(left), _p(right), _p(rate), _p(inv))
        return dict(pos=pos, orient=ori, left=left, right=right, max_rate=rate, inv_max_rate=inv)

    def domain(self):
        d = np.empty(6)
        self.L.t1_domain(self.h, _p(d))
        return d

    def removal_domain(self):
        d = np.empty(6)
        self.L.t1_removal_domain(self.h, _p(d))
        return d

    def inject(self):
        ids = np.empty(self.L.t1_num_inject(self.h), np.int32)
        self.L.t1_inject(self.h, _p(ids))
        return ids

    def bucket_dims(self):
        n = np.empty(3, np.int32)
        self.L.t1_bucket_dims(self.h, _p(n))
        return n

    def row(self, i: int, cap: int = 8192):
        ids, cum = np.empty(cap, np.int32), np.empty(cap)
        d = self.L.t1_row(self.h, i, _p(ids), _p(cum), cap)
        assert d <= cap
        return ids[:d].copy(), cum[:d].copy()

    def csr(self):
        N = self.num_sites()
        deg = np.empty(N, np.int32)
        self.L.t1_degrees(self.h, _p(deg))
        row_ptr = np.zeros(N + 1, np.int64)
        np.cumsum(deg, out=row_ptr[1:])
        ids, cum = np.empty(row_ptr[-1], np.int32), np.empty(row_ptr[-1])
        self.L.t1_csr(self.h, _p(row_ptr), _p(ids), _p(cum))
        return row_ptr, ids, cum

    # -- draws --------------------------------------------------------------------------------------------------
    def draws_glibc(self) -> None:
        self.L.t1_draws_glibc(self.h)

    def draws_philox(self, seed: int) -> None:
        self.L.t1_draws_philox(self.h, seed)

    def draws_replay(self, offsets: np.ndarray, flat: np.ndarray) -> None:
        off = np.ascontiguousarray(offsets, np.int64)
        fl = np.ascontiguousarray(flat, np.int32)
        self._keep = [off, fl]
        self.L.t1_draws_replay(self.h, len(off) - 1, _p(off), _p(fl))

    def log_draws(self, on: bool = True) -> None:
        self.L.t1_log_draws(self.h, 1 if on else 0)

    def trace_sites(self, on: bool = True) -> None:
        self.L.t1_trace_sites(self.h, 1 if on else 0)

    # -- excitons -----------------------------------------------------------------------------------------------
    def create_particles(self, P: int, first_global_id: int = 0) -> None:
        self.L.t1_kubo_create_particles(self.h, P, first_global_id)
        self._n_ids = first_global_id + P

    def kubo_step(self, dt: float, nsteps: int, want_msd: bool = True):
        msd = np.empty((nsteps, 3)) if want_msd else None
        self.L.t1_kubo_step(self.h, dt, nsteps, _p(msd))
        return msd

    def particles(self):
        P = self.L.t1_num_particles(self.h)
        site, heading = np.empty(P, np.int32), np.empty(P, np.int32)
        pos, old, delta, ff = np.empty((3, P)), np.empty((3, P)), np.empty((3, P)), np.empty(P)
        self.L.t1_particles(self.h, _p(site), _p(pos), _p(old), _p(delta), _p(ff), _p(heading))
        return dict(site=site, pos=pos, old_pos=old, delta=delta, ff=ff, heading=heading)

    def time(self) -> float:
        return self.L.t1_time(self.h)

    def hops(self) -> int:
        return self.L.t1_hops(self.h)

    def reinjections(self) -> int:
        return self.L.t1_reinjections(self.h)

    def event_counts(self):
        out = np.empty(self.L.t1_num_particles(self.h), np.int64)
        self.L.t1_event_counts(self.h, _p(out))
        return out

    def _csr_log(self, counts_fn, flat_fn, n_ids):
        counts = np.zeros(n_ids, np.int64)
        counts_fn(self.h, _p(counts))
        off = np.zeros(n_ids + 1, np.int64)
        np.cumsum(counts, out=off[1:])
        flat = np.empty(off[-1], np.int32)
        flat_fn(self.h, _p(flat))
        return off, flat

    def logged_draws(self, n_ids: int):
        return self._csr_log(self.L.t1_draw_counts, self.L.t1_logged_draws, n_ids)

    def traced_sites(self, n_ids: int):
        return self._csr_log(self.L.t1_trace_counts, self.L.t1_traced_sites, n_ids)

    def replay_exhausted(self) -> bool:
        return bool(self.L.t1_replay_exhausted(self.h))

    # -- contacts -----------------------------------------------------------------------------------------------
    def area(self):
        a = np.empty(self.n_seg)
        self.L.t1_area(self.h, _p(a))
        return a

    def contact_sites(self, which: int):
        ids = np.empty(self.L.t1_num_contact_sites(self.h, which), np.int32)
        self.L.t1_contact_sites(self.h, which, _p(ids))
        return ids

    def contact_iteration(self, dt: float):
        pop, cur = np.empty(self.n_seg, np.int64), np.empty(self.n_seg - 1, np.int64)
        self.L.t1_contact_iteration(self.h, dt, _p(pop), _p(cur))
        return pop, cur


    def track_particle(self, dt: float, gid: int = 0, max_steps: int = 1 << 20):
        """monte_carlo::track_particle (monte_carlo.h:786-818): (path [n][3], reached)"""
        path = np.empty((max_steps, 3))
        reached = C.c_int32()
        n = self.L.t1_track_particle(self.h, dt, gid, max_steps, _p(path), C.byref(reached))
        return path[:n].copy(), bool(reached.value)


def load_mc_block(json_path: str) -> dict:
    with open(json_path) as f:
        return json.load(f)["exciton monte carlo"]
