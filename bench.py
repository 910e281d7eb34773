#!/usr/bin/env python
"""Headline benchmark: exciton hops per second on the hop path (BASELINE.json metric).

    python bench.py [--workload C1|C2|C3|C4|C5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads = BASELINE.json's five configs; C2 is the one the metric is quoted on and the default:
  C1  the repo's own input.json (trim limits and dt = 1e-15 s verbatim, 2000 excitons) on the synthetic stand-in for its
      missing mesh (200 tubes x 100 sites); the launch-bound regime
  C2  1000-tube x 100-site random film (seed 1234), 1e6 excitons per GPU, dt = 1e-13 s            (weak scaling)
  C3  the same film, --excitons-total (1e8) split over the ranks by global id                      (strong scaling)
  C4  dense film, 20 000 tubes x 250 sites = 5e6 sites, ~1.4e9 table entries (22 GB): HBM-resident table gathers
  C5  contact-driven transport (monte_carlo::init / step / save_metrics / repopulate_contacts) on the C2 film with the
      contact population scaled to ~1.25e8 excitons alive per GPU (1e9 on 8 GPUs)
One bench "step" = one call of the hop path over all excitons for `--intervals` sampling intervals.  With N GPUs the
per-interval MSD/hop rows (or the integer contact bins) are summed over ranks with one NCCL all-reduce per step.

Prints ONE JSON line (rank 0).  `value` = hops of all ranks / max-over-ranks device time, state resident in HBM;
`e2e` = the same through the host-buffer entry point (cntmc_kubo_step_host_state: H2D + D2H of the whole population
inside the timed region; cntmc_step for the contact flavour).  --impl reference times the reference's own CPU code
(oracle/_ref/libt0.so, OpenMP, all host cores) on a bounded sample of the same workload.
"""
import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DT = 1e-13
TRIM_WIDE = {"xlim": [-1e-5, 1e-5], "ylim": [-1e-5, 1e-5], "zlim": [-1e-5, 1e-5]}
TRIM_INPUT_JSON = {"xlim": [-1e-6, 1e-6], "ylim": [0, 1e-7], "zlim": [-1e-6, 1e-6]}   # the reference's input.json:56-60
WORKLOADS = {
    "C1": dict(film="C1", mode="kubo", dt=1e-15, trim=TRIM_INPUT_JSON, excitons=2000, intervals=20000, chunk=4096, steps=10, scaling="weak",
               text="C1: input.json verbatim (trim limits, dt 1e-15 s, 2000 excitons) on the 200-tube x 100-site stand-in film (seed 1234)"),
    "C2": dict(film="C2", mode="kubo", dt=DT, trim=TRIM_WIDE, excitons=1_000_000, intervals=100, chunk=100, steps=145, scaling="weak",
               text="C2: 1000-tube x 100-site random CNT film (seed 1234), forster table 21x11x11x11, cutoff 20 nm"),
    "C3": dict(film="C2", mode="kubo", dt=DT, trim=TRIM_WIDE, excitons=None, intervals=100, chunk=64, steps=5, scaling="strong",
               text="C3: C2 film, a fixed population (--excitons-total) split over the ranks by global id"),
    "C4": dict(film="C4", mode="kubo", dt=DT, trim=TRIM_WIDE, excitons=1_000_000, intervals=64, chunk=64, steps=10, scaling="weak",
               text="C4: dense film, 20 000 tubes x 250 sites (5e6 sites, ~1.4e9 table entries), HBM-resident rate table"),
    "C5": dict(film="C2", mode="contacts", dt=DT, trim=TRIM_WIDE, excitons=None, intervals=25, chunk=64, steps=5, scaling="weak",
               text="C5: contact-driven transport on the C2 film, contact population scaled to ~1.25e8 excitons alive per GPU"),
    # not a BASELINE config: the step before the hot path, measured to the same contract
    "F1": dict(film=None, mode="davoody", dt=DT, tube=(4, 2, 10), steps=20, scaling="weak",
               grids={"theta [degrees]": [0, 180, 21], "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
                      "axis shift 2 [m]": [-10e-9, 10e-9, 11]},
               text="F1: the davoody rate table of the reference's input.json -- (4,2) tube x 10 unit cells to itself, 21 x 11 x 11 x 11 "
                    "placements, exciton_transfer::first_order per entry"),
}
WORKLOAD = WORKLOADS["C2"]["text"]


def mc_block(P, dt=DT, trim=None):
    return {
        "rate type": "forster",
        "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
        "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21],
        "temperature [kelvin]": 300, "max hopping radius [m]": 20e-9, "number of segments": 10,
        "trim limits": dict(trim or TRIM_WIDE),
        "exciton velocity [m/s]": 2e5, "monte carlo time step": dt,
        "number of sections for injection region": 5,
        "maximum time for kubo simulation [seconds]": 1e-6,
        "number of particles for kubo simulation": P,
    }


def workload_film(name, scale=1.0):
    """(pos, orient) of a workload's film; scale < 1 keeps the density and shrinks the box (CPU samples of C4)."""
    from cnt_film_monte_carlo_b200 import film
    cfg = dict(film.CONFIG_FILMS[WORKLOADS[name]["film"]])
    if scale != 1.0:
        cfg["NT"] = max(8, int(cfg["NT"] * scale))
        cfg["LX"] = cfg["LX"] * scale ** 0.5
    return film.film(**cfg)


def source_hash():
    """Hash of the kernel sources; profiles/ summaries carry the hash of the build they were captured from."""
    h = hashlib.sha1()
    for f in ("kernels.cuh", "hop_core.h", "fast_log.h", "log_table.inc", "csr_core.h"):
        with open(os.path.join(ROOT, "cnt_film_monte_carlo_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def profiled_counters(workload):
    """ncu counters of the hop kernels on this workload from the newest profiles/round*_<workload>_hop_ncu.json
    (written by tools/ncu_to_json.py from one `ncu --set full` capture).  Returns (dict or None, note)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "round*_%s_hop_ncu.json" % workload)))
    if not files:
        return None, "no ncu capture of this workload under profiles/"
    with open(files[-1]) as f:
        prof = json.load(f)
    note = os.path.relpath(files[-1], ROOT)
    if prof.get("source_hash") != source_hash():
        return None, note + " was captured from other kernel sources (hash %s, now %s)" % (prof.get("source_hash"), source_hash())
    return prof, note


class ClockSampler:
    """SM clock and throttle reasons of one GPU, polled through NVML every 25 ms while the timed region runs
    (falls back to `nvidia-smi -lms 100`, the recipe of B200_PROFILING.md, if NVML cannot be loaded)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc, self.stop_flag, self.thread = index, [], None, False, None
        self.sm, self.mx, self.reasons = [], [], set()

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(handle) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                for name, bit in self.BITS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.025)

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = [float(nv.nvmlDeviceGetMaxClockInfo(handle, nv.NVML_CLOCK_SM))]
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        self.stop_flag = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()
            self.sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
            self.mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            self.reasons = {n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)}
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(hops, probes, crossings, P, launches):
    """per hop 16 (row bounds) + 8 per cumulative-rate probe + 4 (neighbour id) + 24 (destination
    position) + 8 (its 1/Gamma), 24 per chain site crossed in flight, plus 2 x 65 B of exciton state per launch."""
    return hops * (16 + 4 + 24 + 8) + 8 * probes + 24 * crossings + 2 * 65 * P * launches


# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's own particle loop on a workload's film, set up once; each step() advances the same simulation
    by `intervals`.  C4's film is sampled at 1/50 of its tubes (same density): the reference needs minutes and tens of
    GB to set up 5e6 sites."""

    def __init__(self, name, P, threads):
        import tempfile
        from cnt_film_monte_carlo_b200 import film
        from oracle import t0 as T0m, t1 as T1m

        wl = WORKLOADS[name]
        self.name, self.dt, self.contacts = name, wl["dt"], wl["mode"] == "contacts"
        self.film_note = "1/50 of the C4 film's tubes at the same density" if name == "C4" else "the workload's film"
        pos, ori = workload_film(name, 0.02 if name == "C4" else 1.0)
        mc = mc_block(P, wl["dt"], wl["trim"])
        self.kind = "reference" if T0m.available() else "port"
        if self.kind == "reference":
            self.tmp = tempfile.TemporaryDirectory()
            mesh = os.path.join(self.tmp.name, "mesh")
            film.write_mesh(mesh, pos, ori)
            mc.update({"mesh input directory": mesh, "output directory": os.path.join(self.tmp.name, "out"), "keep old results": False})
            jpath = os.path.join(self.tmp.name, "input.json")
            with open(jpath, "w") as f:
                json.dump({"exciton monte carlo": mc}, f)
            self.t = T0m.T0()
            self.t.set_threads(threads)
            if self.contacts:
                self.t.open_contacts(jpath, 100)     # monte_carlo::init: contact populations 1100 / 0 (monte_carlo.h:191-192)
            else:
                self.t.open(jpath, 100)
                self.t.create_particles_verbatim()
            self.cores = threads
        else:
            self.t = T1m.T1()
            if self.contacts:
                self.t.draws_glibc()
                self.t.contacts_init(mc, pos, ori)
            else:
                self.t.kubo_init(mc, pos, ori)
                self.t.draws_glibc()
                self.t.create_particles(P)
            self.cores = 1

    def step(self, intervals):
        """(hops, seconds) of `intervals` x kubo_step(dt) (or contact iterations)."""
        if self.contacts:
            if self.kind == "reference":
                d0 = self.t.total_draws()
                t0 = time.perf_counter()
                for _ in range(intervals):
                    self.t.contact_iteration(self.dt)
                sec = time.perf_counter() - t0
                # an event takes two draws, each of the 1100 excitons born per iteration three (site, free flight, heading)
                return max(0, (self.t.total_draws() - d0 - 3 * 1100 * intervals) // 2), sec
            h0 = self.t.hops()
            t0 = time.perf_counter()
            for _ in range(intervals):
                self.t.contact_iteration(self.dt)
            return self.t.hops() - h0, time.perf_counter() - t0
        if self.kind == "reference":
            d0 = self.t.total_draws()
            t0 = time.perf_counter()
            reinj = self.t.kubo_step_omp(self.dt, intervals)
            sec = time.perf_counter() - t0
            return (self.t.total_draws() - d0 - reinj) // 2, sec
        h0 = self.t.hops()
        t0 = time.perf_counter()
        self.t.kubo_step(self.dt, intervals, want_msd=False)
        return self.t.hops() - h0, time.perf_counter() - t0

    def sample_text(self, P, intervals, sec=None):
        what = "contact iterations (contact populations 1100 / 0)" if self.contacts else "%d excitons x" % P
        return "%s %d intervals of %g s on %s%s, reference built against an Armadillo stand-in" % (
            what, intervals, self.dt, self.film_note, "" if sec is None else " (%.1f s)" % sec)


def run_reference(args, rank):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    ref = CpuReference(args.workload, args.cpu_excitons, os.cpu_count() or 1)
    # size the per-step sample so that warm-up + timed steps take about args.cpu_budget seconds in total
    probe = 20 if wl["dt"] >= 1e-14 else 2000
    hops, sec = ref.step(probe)
    per_interval = max(sec / probe, 1e-7)
    intervals = int(min(200000, max(10, args.cpu_budget / per_interval / (args.steps + args.warmup))))
    for _ in range(args.warmup):
        ref.step(intervals)
    vals = [ref.step(intervals) for _ in range(args.steps)]
    hops = sum(v[0] for v in vals)
    sec = sum(v[1] for v in vals)
    value = hops / sec
    sample = ref.sample_text(args.cpu_excitons, intervals)
    print(json.dumps({
        "impl": "reference", "metric": "exciton hops/sec", "value": value, "unit": "hops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, args.steps),
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["text"], "dt_s": wl["dt"], "sample": sample},
        "cpu_baseline": {"value": value, "unit": "hops/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": "hops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(args):
    """The reference timed on this box's host cores on a bounded sample (about 15 s of CPU work)."""
    ref = CpuReference(args.workload, args.cpu_excitons, os.cpu_count() or 1)
    probe = 20 if WORKLOADS[args.workload]["dt"] >= 1e-14 else 2000
    h20, s20 = ref.step(probe)
    cpu_int = int(min(400000, max(probe, 15.0 / max(s20 / probe, 1e-7))))
    hops_c, sec_c = ref.step(cpu_int)
    return {"value": hops_c / sec_c, "unit": "hops/s", "cores": ref.cores, "kind": ref.kind,
            "sample": ref.sample_text(args.cpu_excitons, cpu_int, sec_c)}


# ---------------------------------------------------------------------------------------------------------------------
def dist_setup(local_rank, world):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hop engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    return torch, dist, dev


def timed_steps(torch, dist, dev, stream, world, rank, local_rank, steps, one_step, after_step):
    """K steps bracketed by barrier + synchronize, CUDA events per step on the launching stream, L2 flushed between
    steps (outside the events), max over ranks.  Returns (ms_max, clocks)."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.fill_(0)
        a.record(stream)
        one_step()
        b.record(stream)
        b.synchronize()
        after_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    return float(t_ms.item()), clocks


def run_ours(args, rank, world, local_rank):
    wl = WORKLOADS[args.workload]
    if wl["mode"] == "contacts":
        return run_contacts(args, rank, world, local_rank)
    torch, dist, dev = dist_setup(local_rank, world)
    from cnt_film_monte_carlo_b200.engine import Engine
    from cnt_film_monte_carlo_b200.parallel import shard_range

    dt, n_int = wl["dt"], args.intervals or wl["intervals"]
    if wl["scaling"] == "strong":       # C3: a fixed population split by global id
        first, P = shard_range(args.excitons_total, rank, world)
        P_all = args.excitons_total
    else:                               # weak: every rank holds its own population
        P = args.excitons or wl["excitons"]
        first, P_all = rank * P, P * world
    stream = torch.cuda.current_stream(dev)

    t_setup = time.perf_counter()
    pos, ori = workload_film(args.workload)
    eng = Engine(mc_block(P, dt, wl["trim"]), device=local_rank, stream=stream.cuda_stream)
    eng.set_mesh(pos, ori)
    for k, v in (("chunk_steps", args.chunk or wl["chunk"]), ("hot_pct", args.hot_pct), ("occupancy", args.occupancy), ("top_entries", args.top_entries)):
        eng.set_option(k, v)
    for kv in args.opt:                               # any other engine option, name=value
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    if args.stage_mb > 0:
        eng.set_option("stage_mb", args.stage_mb)
    eng.kubo_init()
    eng.kubo_create_particles(P, seed=1, first_global_id=first)
    t_setup = time.perf_counter() - t_setup

    sums = torch.zeros((n_int, 4), dtype=torch.float64, device=dev)

    def one_step():
        eng.kubo_step_dev(dt, n_int, sums.data_ptr())
        if world > 1:
            dist.all_reduce(sums)   # one NCCL all-reduce of the [intervals][4] MSD / hop histogram per step

    # warm the population up: after the first intervals excitons have left their injection sites
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    acc = {"hops": 0.0, "msd": None}

    def after_step():
        acc["hops"] += float(sums[:, 3].sum().item())   # all ranks' hops after the all-reduce
        acc["msd"] = (sums[-1, :3] / P_all).tolist()

    ms_max, clocks = timed_steps(torch, dist, dev, stream, world, rank, local_rank, args.steps, one_step, after_step)
    hops_total = acc["hops"]
    value = hops_total / (ms_max * 1e-3)

    # ---- dominant kernels: duration measured live (CUDA events around the hop kernels of every launch) + algorithmic bytes
    roof = None
    if rank == 0:
        eng.set_option("time_kernels", 1)
        eng.sync()
        h0 = eng.hops()
        eng.kubo_step(dt, n_int)                       # the kernels that were timed above, now with events around each launch
        k_ms, k_n = eng.kernel_ms(), eng.kernel_launches()
        trap_ms = eng.get_option("dbg_deep_us") / 1e3
        hops = eng.hops() - h0
        step_ms = eng.last_step_ms()
        eng.set_option("time_kernels", 0)
        eng.set_option("stats", 1)                     # instrumented twins: count probes and chain crossings (not timed)
        eng.sync()
        h1, c0, p0 = eng.hops(), eng.crossings(), eng.probes()
        eng.kubo_step(dt, n_int)
        hops_s = max(1, eng.hops() - h1)
        crossings = (eng.crossings() - c0) * hops / hops_s
        probes = (eng.probes() - p0) * hops / hops_s
        walk_share = eng.get_option("dbg_walk_events") / hops_s
        eng.set_option("stats", 0)
        abytes = algorithmic_bytes(hops, probes, crossings, P, k_n)
        peak, which = measured_hbm_peak()
        achieved = abytes / (k_ms * 1e-3) / 1e9
        # DRAM bytes and L2 sectors of the hop kernels from the ncu --set full capture of THIS build on THIS workload
        # (profiles/round*_<workload>_hop_ncu.json carries the hash of the kernel sources it was captured from), scaled
        # from the captured launch to this run's launch size; null when no capture matches the sources
        prof, prof_note = profiled_counters(args.workload)
        traffic = l2 = None
        if prof:
            scale = (P * n_int / max(1, k_n)) / (prof["excitons"] * prof["steps_per_launch"])
            traffic = scale * prof["dram_bytes_per_launch"]
            l2 = {"achieved_gbs_under_ncu": prof["l2_gbs"], "sectors_per_hop": prof["l2_sectors_per_hop"],
                  "dram_gbs_under_ncu": prof["dram_gbs"], "note": "ncu replays are cold-cache and serialised; rates are per kernel"}
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": prof_note, "l2": l2,
                "peak_source": which, "kernel": "kubo_kernel + deep_kernel (trap solver)",
                "kernel_ms_per_launch": k_ms / max(1, k_n), "trap_solver_ms_per_launch": trap_ms / max(1, k_n),
                "kernel_share_of_step": k_ms / step_ms,
                "bytes_per_hop": abytes / max(1, hops), "probes_per_hop": probes / max(1, hops),
                "crossings_per_hop": crossings / max(1, hops), "hops_per_launch": hops / max(1, k_n),
                "events_decided_by_trap_walk": walk_share,
                "note": "HBM lens as the contract asks; tables of C1-C3/C5 are L2-resident and the kernels are bound by dependent-chain latency (DESIGN.md section 4)"}

    # ---- end to end: population in pinned host memory, uploaded and downloaded every step ---------------------------
    state = eng.particles()
    pinned = {}
    for k, v in state.items():
        t = torch.from_numpy(v).pin_memory()
        pinned[k] = t.numpy()
        pinned["_keep_" + k] = t
    h2d = P * (4 + 24 + 24 + 8 + 1 + 4)
    d2h = h2d + n_int * 32
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    rows = torch.zeros((n_int, 4), dtype=torch.float64, device=dev)

    def e2e_step():
        msd = eng.kubo_step_host_state(dt, n_int, pinned)   # synchronous: returns after the D2H
        if world > 1:                                        # the whole-population rows, as in the resident loop
            rows[:, :3].copy_(torch.from_numpy(msd * P))
            dist.all_reduce(rows)

    e2e_step()  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    h_before = eng.hops()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    e2e_local = torch.tensor([float(eng.hops() - h_before), sec], dtype=torch.float64, device=dev)
    if world > 1:
        hh = e2e_local[:1].clone()
        tt = e2e_local[1:].clone()
        dist.all_reduce(hh)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_val = float(hh.item() / tt.item())
    else:
        e2e_val = float(e2e_local[0].item() / e2e_local[1].item())

    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None

    if rank == 0:
        launches_per_step = eng.last_step_launches()
        print(json.dumps({
            "metric": "exciton hops/sec", "value": value, "unit": "hops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": wl["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"], "excitons_per_gpu": P, "excitons_total": P_all, "sites": eng.num_sites(), "dt_s": dt,
                       "intervals_per_step": n_int, "hops_per_step": hops_total / args.steps,
                       "l2": "flushed between timed steps (256 MiB fill, outside the events)",
                       "chunk_steps": args.chunk or wl["chunk"], "hot_pct": args.hot_pct, "occupancy": int(eng.get_option("occupancy")),
                       "options": args.opt, "setup_s": round(t_setup, 2), "table_build_s": eng.csr_build_seconds(),
                       "parallelism": "exciton sharding x%d, tables replicated, 1 all-reduce/step" % world,
                       "msd_last_m2": acc["msd"]},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "hops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_contacts(args, rank, world, local_rank):
    """C5: contact-driven transport.  Every rank holds its share of the contact populations (parallel.ShardedContacts:
    stream ids of different ranks never meet), one NCCL all-reduce of the integer population / current bins per step."""
    wl = WORKLOADS[args.workload]
    torch, dist, dev = dist_setup(local_rank, world)
    from cnt_film_monte_carlo_b200.engine import Engine
    from cnt_film_monte_carlo_b200.parallel import ShardedContacts

    dt, n_int = wl["dt"], args.intervals or wl["intervals"]
    c1_total = args.c1_pop * world
    sc = ShardedContacts(c1_total, 0, 1, rank, world)
    stream = torch.cuda.current_stream(dev)
    t_setup = time.perf_counter()
    pos, ori = workload_film(args.workload)
    eng = Engine(mc_block(1, dt, wl["trim"]), device=local_rank, stream=stream.cuda_stream)
    eng.set_mesh(pos, ori)
    for k, v in (("chunk_steps", args.chunk or wl["chunk"]), ("occupancy", args.occupancy), ("top_entries", args.top_entries)):
        eng.set_option(k, v)
    for kv in args.opt:
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    sc.configure(eng)
    eng.init(sc.c1_pop, sc.c2_pop, seed=sc.seed, capacity=int(7 * sc.c1_pop))
    t_setup = time.perf_counter() - t_setup
    n_seg = eng.number_of_segments()
    bins = torch.zeros((n_int, 2 * n_seg - 1), dtype=torch.int64, device=dev)

    def one_step():
        eng.step_dev(dt, n_int, bins.data_ptr())
        sc.bins(bins)                                   # NCCL all-reduce in place (a few kilobytes)

    for _ in range(args.warmup):
        one_step()
    eng.sync()
    acc = {"h0": eng.hops(), "pop": None}

    def after_step():
        acc["pop"] = int(bins[-1, :n_seg].sum().item())

    ms_max, clocks = timed_steps(torch, dist, dev, stream, world, rank, local_rank, args.steps, one_step, after_step)
    eng.sync()
    cnt = torch.tensor([float(eng.hops() - acc["h0"]), float(eng.number_of_particles())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    hops_total, alive = float(cnt[0].item()), int(cnt[1].item())
    value = hops_total / (ms_max * 1e-3)

    # end to end: the host-buffer entry point (cntmc_step returns the bins of every iteration to host arrays)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    eng.step(dt, n_int)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    h_before = eng.hops()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pop, cur = eng.step(dt, n_int)
        if world > 1:   # whole-ensemble bins, as in the resident loop (NCCL reduces device tensors)
            sc.bins(torch.from_numpy(np.concatenate([pop, cur], axis=1)).to(dev))
    sec = time.perf_counter() - t0
    e2e_local = torch.tensor([float(eng.hops() - h_before), sec], dtype=torch.float64, device=dev)
    if world > 1:
        hh, tt = e2e_local[:1].clone(), e2e_local[1:].clone()
        dist.all_reduce(hh)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_val = float(hh.item() / tt.item())
    else:
        e2e_val = float(e2e_local[0].item() / e2e_local[1].item())

    roof = None
    if rank == 0:
        eng.sync()
        h0, c0, p0 = eng.hops(), eng.crossings(), eng.probes()
        eng.step(dt, n_int)
        ms = eng.last_step_ms()
        hops, crossings, probes = eng.hops() - h0, eng.crossings() - c0, eng.probes() - p0
        launches = max(1, -(-n_int // (args.chunk or wl["chunk"])))
        abytes = algorithmic_bytes(hops, probes, crossings, eng.number_of_particles(), launches)
        peak, which = measured_hbm_peak()
        achieved = abytes / (ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": which, "kernel": "contact_kernel (+ survivor compaction)", "kernel_ms_per_launch": ms / launches,
                "bytes_per_hop": abytes / max(1, hops), "probes_per_hop": probes / max(1, hops), "crossings_per_hop": crossings / max(1, hops)}
    cpu = cpu_baseline(args) if (rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    if rank == 0:
        print(json.dumps({
            "metric": "exciton hops/sec", "value": value, "unit": "hops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"], "c1_pop_per_gpu": sc.c1_pop, "excitons_alive_total": alive, "dt_s": dt,
                       "intervals_per_step": n_int, "hops_per_step": hops_total / args.steps, "population_last_interval": acc["pop"],
                       "l2": "flushed between timed steps (256 MiB fill, outside the events)", "setup_s": round(t_setup, 2),
                       "parallelism": "contact populations split x%d, tables replicated, 1 all-reduce of integer bins/step" % world},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "hops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": n_int * (2 * n_seg - 1) * 8,
                    "steps": e2e_steps, "note": "the contact population is created and destroyed on the device; the host receives the bins"},
            "gpu_launches": int(eng.last_step_launches() * args.steps),
            "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ---- F1: the davoody rate table -------------------------------------------------------------------------------------------
def davoody_reference_sample(wl, n_theta):
    """The reference's own first_order (oracle/_ref/libf1.so, its loop nest over theta under OpenMP like monte_carlo.cpp:107-137) on
    the first n_theta angles x 1 z shift x 2 x 2 axis shifts of the workload's table.  Returns (entries, seconds, threads)."""
    from cnt_film_monte_carlo_b200 import davoody as dv
    from oracle import f1
    if not f1.available():
        return None
    theta, z, a1, a2 = dv.table_axes(wl["grids"])
    r = f1.RefTube(*wl["tube"])
    f1.table(r, r, theta[:1], z[:1], a1[:1], a2[:1])  # first call builds the matched state pairs
    t0 = time.perf_counter()
    out = f1.table(r, r, theta[:n_theta], z[:1], a1[4:6], a2[4:6])
    return out.size, time.perf_counter() - t0, min(n_theta, os.cpu_count() or 1)


def run_davoody(args, rank, world, local_rank):
    """One step = the whole table of the workload through the C ABI (cntmc_transfer_table: host axes in, host rates out).  `value`
    counts the placement kernel alone (CUDA events around it on its stream: the placements are in device memory by then), `e2e`
    the whole call.  Ranks build replicas (the entries are independent; nothing to exchange)."""
    wl = WORKLOADS["F1"]
    if args.impl == "reference":
        if rank != 0:
            return
        n_theta = min(21, max(2, os.cpu_count() or 1))
        got = davoody_reference_sample(wl, n_theta)
        if got is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libf1.so is not built on this box"}))
            return
        vals = [davoody_reference_sample(wl, n_theta) for _ in range(max(1, min(args.steps, 5)))]
        n, sec = sum(v[0] for v in vals), sum(v[1] for v in vals)
        sample = "%d table entries per step (%d angles x 1 x 2 x 2), reference first_order under OpenMP over theta" % (vals[0][0], n_theta)
        print(json.dumps({
            "impl": "reference", "metric": "davoody table entries/sec", "value": n / sec, "unit": "entries/s", "n_gpus": args.gpus,
            "steps": len(vals), "warmup": 1, "ms_per_step": 1e3 * sec / len(vals), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": wl["text"], "sample": sample},
            "cpu_baseline": {"value": n / sec, "unit": "entries/s", "cores": vals[0][2], "kind": "reference", "sample": sample},
            "e2e": {"value": n / sec, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    torch, dist, dev = dist_setup(local_rank, world)
    from cnt_film_monte_carlo_b200 import davoody as dv

    t0 = time.perf_counter()
    tube = dv.Tube(*wl["tube"])
    x = dv.Transfer(tube, tube, device=local_rank)
    t_setup = time.perf_counter() - t0
    axes = dv.table_axes(wl["grids"])
    n_entries = int(np.prod([len(a) for a in axes]))
    info = x.info()
    passes = -(-info["donor_kcm"] // info["kcm_per_pass"])
    # FP64 operations the formula needs per site pair and pass: 3 sub, 3 mul, 2 add, sqrt, div + 2 fused multiply-adds per K_cm
    flop_per_step = float(n_entries) * tube.sites * tube.sites * passes * (10 + 4 * info["kcm_per_pass"])
    for _ in range(max(3, args.warmup)):
        x.table(*axes)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    kernel_ms, wall = 0.0, 0.0
    for _ in range(args.steps):
        flush.fill_(0)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        rates = x.table(*axes)       # synchronous: returns after the D2H of the rates
        wall += time.perf_counter() - w0
        kernel_ms += x.info()["last_kernel_ms"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t = torch.tensor([kernel_ms, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    kernel_ms, wall = float(t[0].item()), float(t[1].item())
    value = world * n_entries * args.steps / (kernel_ms * 1e-3)
    achieved = flop_per_step * args.steps / (kernel_ms * 1e-3) / 1e12
    peak = dv.fp64_peak(local_rank)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        got = davoody_reference_sample(wl, min(21, max(2, os.cpu_count() or 1)))
        if got is not None:
            cpu = {"value": got[0] / got[1], "unit": "entries/s", "cores": got[2], "kind": "reference",
                   "sample": "%d table entries (%d angles x 1 x 2 x 2) in %.1f s" % (got[0], got[2], got[1])}
    if rank == 0:
        print(json.dumps({
            "metric": "davoody table entries/sec", "value": value, "unit": "entries/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": kernel_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["text"], "entries_per_step": n_entries, "sites_per_tube": tube.sites, "state_pairs": info["pairs"],
                       "donor_kcm": info["donor_kcm"], "kcm_per_pass": info["kcm_per_pass"], "threads_per_block": info["threads"],
                       "tube_and_transfer_setup_s": round(t_setup, 3), "rate_min_max": [float(rates.min()), float(rates.max())],
                       "l2": "flushed between timed steps (256 MiB fill, outside the timed calls)",
                       "parallelism": "replicas x%d (independent table entries, no exchange)" % world},
            "clocks": clocks,
            "e2e": {"value": world * n_entries * args.steps / wall, "unit": "entries/s", "h2d_bytes_per_step": 5 * 8 * n_entries,
                    "d2h_bytes_per_step": 8 * n_entries, "steps": args.steps},
            "gpu_launches": args.steps,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                         "peak_source": "measured on this GPU (cntmc_fp64_peak: register-only fused multiply-add kernel)",
                         "kernel": "placement_rate_kernel<%d>" % info["kcm_per_pass"], "kernel_ms_per_launch": kernel_ms / args.steps,
                         "note": "counted: 10 + 4 K FP64 operations per site pair; a correctly rounded 1/sqrt is ~24 of the ~40 FP64 "
                                 "instructions per pair and counts as 2 (ncu: FP64 pipe 66 % busy, profiles/round2_davoody_*)"},
            "cpu_baseline": cpu}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0,
                    help="timed bench steps (0 = the workload's own; C2: 145 x 100 intervals = the config's 1e4 hops per exciton)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS), help="BASELINE.json config (C2 = the headline)")
    ap.add_argument("--excitons", type=int, default=0, help="excitons per GPU (0 = the workload's own)")
    ap.add_argument("--excitons-total", type=int, default=100_000_000, help="C3: whole population, split over the ranks")
    ap.add_argument("--c1-pop", type=int, default=22_000_000, help="C5: contact-1 population per GPU (~5.7x as many excitons alive)")
    ap.add_argument("--intervals", type=int, default=0, help="sampling intervals per bench step (0 = the workload's own)")
    ap.add_argument("--chunk", type=int, default=0, help="time steps per kernel launch (0 = the workload's own)")
    ap.add_argument("--hot-pct", type=int, default=25, help="share of blocks serving the most active excitons first")
    ap.add_argument("--opt", action="append", default=[], help="extra engine option name=value (repeatable)")
    ap.add_argument("--top-entries", type=int, default=1, help="1: the three widest entries of a row are tried before the row is searched")
    ap.add_argument("--occupancy", type=int, default=0, help="resident 128-thread blocks per SM of the hop kernel (4 to 8; 0 = the engine's choice: 8 when the tables exceed L2 or for 2e6 excitons and more, else 7)")
    ap.add_argument("--stage-mb", type=int, default=0, help="cap on the (step, exciton) staging buffer in MiB (0 = engine default)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-excitons", type=int, default=8000)
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="seconds of CPU work for the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = WORKLOADS[args.workload]["steps"]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if WORKLOADS[args.workload]["mode"] == "davoody":
        run_davoody(args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
