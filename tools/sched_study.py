"""Warp-scheduling study on the CPU: how many warp-instructions do different loop structures of the hop kernel need?

Uses the host-compiled engine core (tests/host_emul) to record, per exciton, the exact sequence of micro-operations of
one launch on the C2 film, packs excitons into warps of 32 (in index order or sorted by the previous launch's event
count) and replays three loop structures under a simple instruction-cost model.  Not part of the product.
"""
import ctypes as C
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from emul import Emul, _p
from cnt_film_monte_carlo_b200 import film
from conftest import base_mc

COST_C, COST_MOVE, COST_E, COST_S, COST_RED = 80, 60, 250, 50, 60

def sequences(P=32768, nsteps=64, warm=64):
    e = Emul(base_mc()); e.kubo_init(*film.film(**film.CONFIG_FILMS["C2"]))
    e.create_philox(P, seed=1)
    e.L.emul_op_sequences.restype = C.c_int64
    e.L.emul_op_sequences.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
    out = []
    for n in (warm, nsteps):
        counts = np.zeros(P, np.int64); cap = 400 * P * max(1, n // 16)
        flat = np.zeros(cap, np.int32)
        k = e.L.emul_op_sequences(e.h, 1e-13, n, _p(counts), _p(flat), cap)
        assert k <= cap
        off = np.zeros(P + 1, np.int64); np.cumsum(counts, out=off[1:])
        out.append((off, flat[:k]))
    return out

def simulate(off, ops, order, nsteps):
    P = len(off) - 1
    tot = {"P0": 0, "P1": 0, "P2": 0}; ideal = 0
    for w in range(0, P, 32):
        lanes = order[w:w + 32]
        seqs = [ops[off[i]:off[i + 1]] for i in lanes]
        # ideal: every lane-op at full occupancy
        for s in seqs:
            cr = (s >> 1).sum(); ns = (s & 1).sum(); ne = len(s) - ns
            ideal += (cr * COST_C + len(s) * COST_MOVE + ne * COST_E + ns * COST_S) / 32.0
        # P0: one advance() per iteration
        L = max(len(s) for s in seqs)
        pos_step = np.zeros(len(seqs), int)
        for it in range(L):
            act = [s[it] for s in seqs if it < len(s)]
            mc = max(a >> 1 for a in act)
            anyE = any((a & 1) == 0 for a in act); anyS = any(a & 1 for a in act)
            tot["P0"] += mc * COST_C + COST_C + COST_MOVE + (COST_E if anyE else 0) + (COST_S + COST_RED if anyS else 0)
        # micro-op streams: each advance = k crossings 'C' then terminal 'E'/'S' (the terminal includes the failed attempt)
        streams = []
        for s in seqs:
            st = []
            for a in s:
                st.extend([0] * (a >> 1)); st.append(2 if (a & 1) else 1)
            streams.append(st)
        # P1: every iteration runs the common flight attempt, then E part if any lane needs it, then S part
        L = max(len(s) for s in streams)
        for it in range(L):
            act = [s[it] for s in streams if it < len(s)]
            tot["P1"] += COST_C + (COST_MOVE if any(a for a in act) else 0) + (COST_E if 1 in act else 0) + (COST_S if 2 in act else 0)
        # P2: flight attempts always run for lanes in flight; a lane whose flight ended parks with a pending E or S;
        #     pending terminals run when >= half of the unfinished lanes wait for that kind (or nobody can fly)
        idx = [0] * len(streams); pend = [None] * len(streams)
        while True:
            alive = [k for k in range(len(streams)) if idx[k] < len(streams[k]) or pend[k] is not None]
            if not alive: break
            fly = [k for k in alive if pend[k] is None]
            nE = sum(1 for k in alive if pend[k] == 1); nS = sum(1 for k in alive if pend[k] == 2)
            if fly and max(nE, nS) < 0.5 * len(alive):
                tot["P2"] += COST_C + COST_MOVE * 0.3
                for k in fly:
                    op = streams[k][idx[k]]; idx[k] += 1
                    if op: pend[k] = op
            else:
                kind = 1 if (nE * COST_E >= nS * COST_S and nE) or not nS else 2
                tot["P2"] += COST_MOVE + (COST_E if kind == 1 else COST_S)
                for k in alive:
                    if pend[k] == kind: pend[k] = None
    return ideal, tot

if __name__ == "__main__":
    (off0, ops0), (off1, ops1) = sequences()
    P = len(off1) - 1
    ev_prev = np.array([np.count_nonzero((ops0[off0[i]:off0[i + 1]] & 1) == 0) for i in range(P)])
    ev_now = np.array([np.count_nonzero((ops1[off1[i]:off1[i + 1]] & 1) == 0) for i in range(P)])
    print("events/exciton this launch: mean %.1f p50 %d p90 %d p99 %d max %d; corr(prev, now) = %.3f" % (
        ev_now.mean(), np.median(ev_now), np.percentile(ev_now, 90), np.percentile(ev_now, 99), ev_now.max(), np.corrcoef(ev_prev, ev_now)[0, 1]))
    for name, order in (("index order", np.arange(P)), ("sorted by previous events", np.argsort(-ev_prev, kind="stable")),
                        ("sorted by oracle (this launch)", np.argsort(-ev_now, kind="stable"))):
        ideal, tot = simulate(off1, ops1, order, 64)
        print("%-32s ideal %.3e | " % (name, ideal) + " | ".join("%s %.3e (eff %.0f%%)" % (k, v, 100 * ideal / v) for k, v in tot.items()))
