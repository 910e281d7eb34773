"""Quick look at the two-engine hop kernel on the bench workload: ms per 64-step launch, lane-iteration counters."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
import bench

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
variants = sys.argv[2:] or ["engines=0,fast_rounds=0", "engines=1,fast_rounds=8"]
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
for v in variants:
    opts = dict((k, int(x)) for k, x in (kv.split("=") for kv in v.split(",")))
    e = Engine(bench.mc_block(P), device=0)
    e.set_mesh(pos, ori)
    for k, x in opts.items():
        e.set_option(k, x)
    e.kubo_init()
    e.kubo_create_particles(P, seed=1)
    e.kubo_step(1e-13, 64, want_msd=False)
    t = []
    for _ in range(3):
        h0 = e.hops()
        e.kubo_step(1e-13, 64, want_msd=False)
        t.append((e.last_step_ms(), e.hops() - h0))
    e.set_option("stats", 1)
    h0 = e.hops()
    e.kubo_step(1e-13, 64, want_msd=False)
    g = lambda n: e.L.cntmc_get_option(e.h, n.encode())
    print(v, "ms/launch", [round(a, 3) for a, _ in t], "hops/s %.3e" % (t[-1][1] / t[-1][0] * 1e3),
          "| instr: ms", round(e.last_step_ms(), 3), "hops", e.hops() - h0, "fast", g("dbg_fast_events"), "busy", g("dbg_lane_busy"), "idle", g("dbg_lane_idle"),
          "warps", g("dbg_warps"), "span_ms", g("dbg_span_ns") / 1e6, "mean_warp_ms", g("dbg_warp_ns") / max(1, g("dbg_warps")) / 1e6, flush=True)
    if opts.get("engines", 1):
        W = max(1, g("dbg_warps"))
        for v, name in ((0, "flight"), (1, "event ")):
            ph = [g("dbg_phase_%d%d" % (v, k)) / W / 1e6 for k in range(6)]
            it = g("dbg_phase_it%d" % v)
            print("   %s engine: ms per warp (mean over all warps) refill %.3f  step-ends %.3f  fast %.3f  events %.3f  hand-over %.3f  waiting %.3f | iterations %d  (us per iteration %.2f)"
                  % (name, *ph, it, sum(ph[:5]) * W * 1e3 / max(1, it)), flush=True)
