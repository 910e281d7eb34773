"""Scheduling study, part 5: several operations per lane per iteration.  The loop body becomes a fixed sequence of
phases (e.g. S,E or E,S,E); in each phase every lane whose next operation has that kind executes it."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_study import sequences
from sched_study4 import ops_of, C_X, C_E, C_S

def run_warp(queue, off, ops, phases):
    q = list(queue); cost = 0.0; ideal = 0.0
    cur = [None] * 32; idx = [0] * 32
    def refill(k):
        nonlocal ideal
        if q:
            cur[k] = ops_of(off, ops, q.pop(0)); idx[k] = 0
            for cr, se in cur[k]: ideal += (cr * C_X + (C_S if se else C_E)) / 32.0
        else: cur[k] = None
    for k in range(32): refill(k)
    while any(c is not None for c in cur):
        cost += 30
        for ph in phases:            # ph: 1 = step-end kind, 0 = event kind
            run = [k for k in range(32) if cur[k] is not None and cur[k][idx[k]][1] == ph]
            if not run: 
                cost += 4; continue
            mc = max(cur[k][idx[k]][0] for k in run)
            cost += mc * C_X + (C_S if ph else C_E)
            for k in run:
                idx[k] += 1
                if idx[k] >= len(cur[k]): refill(k)   # refill only happens after a step end (last op of an exciton)
    return cost, ideal

if __name__ == "__main__":
    (off0, ops0), (off1, ops1) = sequences(P=16384)
    P = len(off1) - 1
    ev_prev = np.array([np.count_nonzero((ops0[off0[i]:off0[i + 1]] & 1) == 0) for i in range(P)])
    order = np.argsort(-ev_prev, kind="stable")
    W_ = 64
    for phases in ((1, 0), (0, 1), (0, 1, 0), (0, 0, 1), (1, 0, 0), (0, 0, 0, 1), (0, 0, 1, 0, 0)):
        res = [run_warp(order[w::W_], off1, ops1, phases) for w in range(W_)]
        c = np.array([r[0] for r in res]); i = np.array([r[1] for r in res])
        print("phases %-16s eff %.0f%%  total %.3e  makespan/mean %.2f" % ("".join("SE"[1 - p] if False else ("S" if p else "E") for p in phases), 100 * i.sum() / c.sum(), c.sum(), c.max() / c.mean()))
