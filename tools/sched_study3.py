"""Scheduling study, part 3: two-phase launch.  Phase A gives every exciton an event budget; an exciton that exhausts it
is parked and finished by phase B, whose warps therefore hold event-heavy lanes only.  Both phases refill lanes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_study import sequences, COST_C, COST_MOVE, COST_E, COST_S
import sched_study2 as s2

def split(off, ops, budget):
    """Return per-exciton (streamA, streamB) micro-op lists."""
    A, B = [], []
    for i in range(len(off) - 1):
        st = s2.streams_of(off, ops, i)
        ev = 0; cut = len(st)
        for k, op in enumerate(st):
            if op == 1:
                ev += 1
                if ev >= budget: cut = k + 1; break
        A.append(st[:cut]); B.append(st[cut:])
    return A, B

def run(streams, order, W, policy="P0"):
    # reuse run_warp by faking off/ops through a closure
    tot = idl = 0.0; per = []
    old = s2.streams_of
    try:
        s2.streams_of = lambda off, ops, i: streams[i]
        for w in range(W):
            c, d = s2.run_warp(order[w::W], None, None, policy)
            tot += c; idl += d; per.append(c)
    finally:
        s2.streams_of = old
    return tot, idl, (max(per) if per else 0), (np.mean(per) if per else 0)

if __name__ == "__main__":
    (off0, ops0), (off1, ops1) = sequences(P=16384)
    P = len(off1) - 1
    ev_prev = np.array([np.count_nonzero((ops0[off0[i]:off0[i + 1]] & 1) == 0) for i in range(P)])
    base_order = np.argsort(-ev_prev, kind="stable")
    for budget in (8, 16, 32, 64):
        A, B = split(off1, ops1, budget)
        hot = np.array([i for i in range(P) if B[i]])
        ta, ia, ma, mea = run(A, base_order, 64)
        hot_order = hot[np.argsort(-np.array([len(B[i]) for i in hot]), kind="stable")]  # oracle LPT inside B: optimistic
        hot_order_pred = hot[np.argsort(-ev_prev[hot], kind="stable")]
        wb = max(1, min(64, len(hot) // 64))
        tb, ib, mb, meb = run(B, hot_order_pred, wb)
        print("budget %3d: hot %5d (%.1f%%) | A total %.3e eff %.0f%% | B total %.3e eff %.0f%% (warps %d, makespan/mean %.2f) | overall eff %.0f%%" % (
            budget, len(hot), 100 * len(hot) / P, ta, 100 * ia / ta, tb, 100 * ib / max(tb, 1), wb, mb / max(meb, 1), 100 * (ia + ib) / (ta + tb)))
