"""Scheduling study, part 4 (cost model of the v3 core): persistent warps with refill; policy P0 = run whatever each lane
needs every iteration; P3 = lanes whose next operation is an event park until enough of them wait (or one has waited W
iterations).  Costs in warp-instructions: crossing 12, event tail 340, step-end tail 200."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_study import sequences
C_X, C_E, C_S = 12, 340, 200

def ops_of(off, ops, i):
    a = ops[off[i]:off[i + 1]]
    return list(zip((a >> 1).tolist(), (a & 1).tolist()))   # (crossings, is_step_end)

def run_warp(queue, off, ops, policy, theta=0.25, W=6):
    q = list(queue); cost = 0.0; ideal = 0.0; iters = 0
    cur = [None] * 32; idx = [0] * 32; wait = [0] * 32
    def refill(k):
        nonlocal ideal
        if q:
            cur[k] = ops_of(off, ops, q.pop(0)); idx[k] = 0; wait[k] = 0
            for cr, se in cur[k]: ideal += (cr * C_X + (C_S if se else C_E)) / 32.0
        else: cur[k] = None
    for k in range(32): refill(k)
    while True:
        alive = [k for k in range(32) if cur[k] is not None]
        if not alive: break
        iters += 1
        nxt = {k: cur[k][idx[k]] for k in alive}
        if policy == "P0":
            run = alive
        else:
            E = [k for k in alive if nxt[k][1] == 0]; S = [k for k in alive if nxt[k][1] == 1]
            fire = (not S) or len(E) >= theta * len(alive) or any(wait[k] >= W for k in E)
            run = S + (E if fire else [])
            for k in E: wait[k] = 0 if fire else wait[k] + 1
        mc = max(nxt[k][0] for k in run)
        cost += mc * C_X + (C_E if any(nxt[k][1] == 0 for k in run) else 0) + (C_S if any(nxt[k][1] == 1 for k in run) else 0) + 30
        for k in run:
            idx[k] += 1
            if idx[k] >= len(cur[k]): refill(k)
    return cost, ideal, iters

if __name__ == "__main__":
    (off0, ops0), (off1, ops1) = sequences(P=16384)
    P = len(off1) - 1
    ev_prev = np.array([np.count_nonzero((ops0[off0[i]:off0[i + 1]] & 1) == 0) for i in range(P)])
    order = np.argsort(-ev_prev, kind="stable")
    W_ = 64
    for name, pol, kw in (("P0", "P0", {}), ("P3 th.25 W6", "P3", dict(theta=0.25, W=6)), ("P3 th.25 W3", "P3", dict(theta=0.25, W=3)),
                          ("P3 th.5 W8", "P3", dict(theta=0.5, W=8)), ("P3 th.15 W4", "P3", dict(theta=0.15, W=4))):
        for qname, mk in (("interleaved", lambda w: order[w::W_]), ("blocked (hot warps / cold warps)", lambda w: order[w * (P // W_):(w + 1) * (P // W_)])):
            res = [run_warp(mk(w), off1, ops1, pol, **kw) for w in range(W_)]
            c = np.array([r[0] for r in res]); i = np.array([r[1] for r in res])
            print("%-14s %-34s eff %.0f%%  total %.3e  makespan/mean %.2f" % (name, qname, 100 * i.sum() / c.sum(), c.sum(), c.max() / c.mean()))
