"""Warp residency of one instrumented hop-kernel launch on the bench workload: how long the kernel ran (device clock),
how long the average warp stayed, and how many lane-iterations had no exciton (drain).   python tools/residency.py [P] [chunk] [hot_pct]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from bench import mc_block, DT
P = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 64
hot = int(sys.argv[3]) if len(sys.argv) > 3 else 30
pos, ori = film.film(**film.CONFIG_FILMS["C2"])
e = Engine(mc_block(P)); e.set_mesh(pos, ori); e.set_option("chunk_steps", chunk); e.set_option("hot_pct", hot)
e.kubo_init(); e.kubo_create_particles(P, seed=1)
e.kubo_step(DT, 320, want_msd=False)
e.kubo_step(DT, chunk, want_msd=False); plain_ms = e.last_step_ms()
e.set_option("stats", 1)
e.kubo_step(DT, chunk, want_msd=False)
g = lambda k: e.get_option(k)
span, warp_ns, warps, busy, idle = g("dbg_span_ns"), g("dbg_warp_ns"), g("dbg_warps"), g("dbg_lane_busy"), g("dbg_lane_idle")
print(json.dumps({"P": P, "chunk": chunk, "hot_pct": hot, "plain_call_ms": plain_ms, "instr_call_ms": e.last_step_ms(), "kernel_span_ms": span / 1e6,
                  "mean_warp_residency_ms": warp_ns / max(1, warps) / 1e6, "warps": warps, "warp_time_utilisation": warp_ns / max(1, warps) / max(1, span),
                  "lane_iterations_busy": busy, "lane_iterations_idle": idle, "idle_share": idle / max(1, busy + idle)}))
