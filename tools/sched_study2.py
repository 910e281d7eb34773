"""Scheduling study, part 2: persistent warps whose lanes refill from a queue ordered by predicted activity."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sched_study import sequences, COST_C, COST_MOVE, COST_E, COST_S

def streams_of(off, ops, i):
    st = []
    for a in ops[off[i]:off[i + 1]]:
        st.extend([0] * (a >> 1)); st.append(2 if (a & 1) else 1)
    return st

def run_warp(queue, off, ops, policy):
    """queue: exciton ids for this warp, pulled in order by lanes as they free up.  Returns (cost, ideal)."""
    q = list(queue); cost = 0.0; ideal = 0.0
    cur = [None] * 32; idx = [0] * 32; pend = [0] * 32
    def refill(k):
        if q:
            cur[k] = streams_of(off, ops, q.pop(0)); idx[k] = 0; pend[k] = 0
            s = np.array(cur[k]); ideal_add = ((s == 0).sum() * COST_C + (s > 0).sum() * COST_MOVE + (s == 1).sum() * COST_E + (s == 2).sum() * COST_S) / 32.0
            return ideal_add
        cur[k] = None; return 0.0
    for k in range(32): ideal += refill(k)
    while True:
        alive = [k for k in range(32) if cur[k] is not None]
        if not alive: break
        if policy == "P1":   # micro-op per iteration: common flight attempt, then E part, then S part
            acts = []
            for k in alive:
                acts.append(cur[k][idx[k]]); idx[k] += 1
            cost += COST_C + (COST_MOVE if any(acts) else 0) + (COST_E if 1 in acts else 0) + (COST_S if 2 in acts else 0)
        elif policy == "P0":  # one advance() per iteration: flight loop to the end, then E / S
            mc = 0; anyE = anyS = False
            for k in alive:
                c = 0
                while cur[k][idx[k]] == 0: c += 1; idx[k] += 1
                t = cur[k][idx[k]]; idx[k] += 1
                mc = max(mc, c); anyE |= t == 1; anyS |= t == 2
            cost += (mc + 1) * COST_C + COST_MOVE + (COST_E if anyE else 0) + (COST_S if anyS else 0)
        else:                 # P2: park lanes at their terminal; run a terminal kind when >= frac of lanes wait for it
            fly = [k for k in alive if pend[k] == 0]
            nE = sum(1 for k in alive if pend[k] == 1); nS = sum(1 for k in alive if pend[k] == 2)
            thr = policy[1]
            if fly and nE < thr[0] * len(alive) and nS < thr[1] * len(alive):
                cost += COST_C + 0.3 * COST_MOVE
                for k in fly:
                    op = cur[k][idx[k]]; idx[k] += 1
                    if op: pend[k] = op
            else:
                kind = 1 if (nE and (nE >= thr[0] * len(alive) or not nS or not fly and nE * COST_E >= nS * COST_S)) else 2
                if kind == 2 and nS == 0: kind = 1
                cost += COST_MOVE + (COST_E if kind == 1 else COST_S)
                for k in alive:
                    if pend[k] == kind: pend[k] = 0
        for k in alive:
            if pend[k] == 0 and idx[k] >= len(cur[k]): ideal += refill(k)
    return cost, ideal

if __name__ == "__main__":
    (off0, ops0), (off1, ops1) = sequences(P=16384)
    P = len(off1) - 1
    ev_prev = np.array([np.count_nonzero((ops0[off0[i]:off0[i + 1]] & 1) == 0) for i in range(P)])
    order = np.argsort(-ev_prev, kind="stable")
    W = 64
    for policy in ("P0", "P1", ("P2", (0.5, 0.5)), ("P2", (0.25, 0.5)), ("P2", (0.25, 0.25))):
        costs, ideals = zip(*[run_warp(order[w::W], off1, ops1, policy) for w in range(W)])
        print(policy, "total %.3e ideal %.3e eff %.0f%% | makespan/mean %.2f" % (sum(costs), sum(ideals), 100 * sum(ideals) / sum(costs), max(costs) / np.mean(costs)))
