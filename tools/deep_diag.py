"""Where the time of a launch goes with the trap solver on: lane kernel vs trap solver, list sizes, share of the events
the window walk decides.  python tools/deep_diag.py [P] [deep_thr] [name=value ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 16
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(P)); e.set_mesh(pos, ori); e.set_option("deep_thr", thr)
for kv in sys.argv[3:]:
    k, v = kv.split("="); e.set_option(k, int(v))
e.kubo_init(); e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 300, want_msd=False)
out = {"P": P, "deep_thr": thr}
e.set_option("time_kernels", 1)
h0 = e.hops(); e.kubo_step(DT, 64, want_msd=False)
out.update(hops=e.hops() - h0, kernels_ms=e.kernel_ms(), deep_ms=e.get_option("dbg_deep_us") / 1e3, step_ms=e.last_step_ms(),
           class4=e.get_option("dbg_class4"), deferred_last_round=e.get_option("dbg_deferred"), returned=e.get_option("dbg_returned"))
e.set_option("time_kernels", 0); e.set_option("stats", 1)
h0 = e.hops(); e.kubo_step(DT, 64, want_msd=False)
out.update(hops_instr=e.hops() - h0, top_events=e.get_option("dbg_top_events"), walk_events=e.get_option("dbg_walk_events"),
           lane_busy=e.get_option("dbg_lane_busy"), lane_idle=e.get_option("dbg_lane_idle"), warps=e.get_option("dbg_warps"),
           warp_ns=e.get_option("dbg_warp_ns"), span_ns=e.get_option("dbg_span_ns"))
print(json.dumps(out))
