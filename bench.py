#!/usr/bin/env python
"""Headline benchmark: exciton hops per second on the hop path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY.md §8d "C2"): synthetic 1000-tube x 100-site CNT film (seed 1234), closed-
form Forster table on the input.json grids, 20 nm cut-off, 1e6 excitons per GPU, sampling interval dt = 1e-13 s.
One bench "step" = one call of the hop path over all excitons for `--intervals` sampling intervals (default 100).
With N GPUs every rank holds its own 1e6 excitons (weak scaling; exciton streams are keyed by global id) and the
per-interval MSD/hop histogram is summed over ranks with one NCCL all-reduce per step.

Prints ONE JSON line (rank 0).  `value` = hops of all ranks / max-over-ranks device time, state resident in HBM;
`e2e` = the same through cntmc_kubo_step_host_state with the exciton population in pinned HOST memory (H2D + D2H of the
whole state inside the timed region).  --impl reference times the reference's own CPU code (oracle/_ref/libt0.so,
OpenMP, all host cores) on a bounded sample of the same workload.
"""
import argparse
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
WORKLOAD = "C2: 1000-tube x 100-site random CNT film (seed 1234), forster table 21x11x11x11, cutoff 20 nm"


def mc_block(P):
    return {
        "rate type": "forster",
        "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
        "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21],
        "temperature [kelvin]": 300, "max hopping radius [m]": 20e-9, "number of segments": 10,
        "trim limits": {"xlim": [-1e-5, 1e-5], "ylim": [-1e-5, 1e-5], "zlim": [-1e-5, 1e-5]},
        "exciton velocity [m/s]": 2e5, "monte carlo time step": DT,
        "number of sections for injection region": 5,
        "maximum time for kubo simulation [seconds]": 1e-6,
        "number of particles for kubo simulation": P,
    }


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


NCU_DRAM_BYTES_PER_EXCITON_STEP = (373.112320e6 + 1.285876e9) / (36 * 1_000_000)


def algorithmic_bytes(hops, probes, crossings, P, launches):
    """per hop 16 (row bounds) + 8 per cumulative-rate probe + 4 (neighbour id) + 24 (destination
    position) + 8 (its 1/Gamma), 24 per chain site crossed in flight, plus 2 x 65 B of exciton state per launch."""
    return hops * (16 + 4 + 24 + 8) + 8 * probes + 24 * crossings + 2 * 65 * P * launches


# ---------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's own particle loop, set up once; each step() advances the same simulation by `intervals`."""

    def __init__(self, P, threads):
        import tempfile
        from cnt_film_monte_carlo_b200 import film
        from oracle import t0 as T0m, t1 as T1m

        pos, ori = film.film(**film.CONFIG_FILMS["C2"])
        mc = mc_block(P)
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
            self.t.open(jpath, 100)
            self.t.create_particles_verbatim()
            self.cores = threads
        else:
            self.t = T1m.T1()
            self.t.kubo_init(mc, pos, ori)
            self.t.draws_glibc()
            self.t.create_particles(P)
            self.cores = 1

    def step(self, intervals):
        """(hops, seconds) of `intervals` x kubo_step(1e-13)."""
        if self.kind == "reference":
            d0 = self.t.total_draws()
            t0 = time.perf_counter()
            reinj = self.t.kubo_step_omp(DT, intervals)
            sec = time.perf_counter() - t0
            return (self.t.total_draws() - d0 - reinj) // 2, sec
        h0 = self.t.hops()
        t0 = time.perf_counter()
        self.t.kubo_step(DT, intervals, want_msd=False)
        return self.t.hops() - h0, time.perf_counter() - t0


def run_reference(args, rank):
    if rank != 0:
        return
    ref = CpuReference(args.cpu_excitons, os.cpu_count() or 1)
    # size the per-step sample so that warm-up + timed steps take about args.cpu_budget seconds in total
    hops, sec = ref.step(20)
    rate = max(hops / sec, 1.0)
    per_interval = max(sec / 20, 1e-6)
    intervals = int(min(2000, max(10, args.cpu_budget / per_interval / (args.steps + args.warmup))))
    for _ in range(args.warmup):
        ref.step(intervals)
    vals = [ref.step(intervals) for _ in range(args.steps)]
    hops = sum(v[0] for v in vals)
    sec = sum(v[1] for v in vals)
    value = hops / sec
    sample = "%d excitons x %d intervals of 1e-13 s per step on the C2 film, reference built against an Armadillo stand-in" % (
        args.cpu_excitons, intervals)
    print(json.dumps({
        "impl": "reference", "metric": "exciton hops/sec", "value": value, "unit": "hops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "dt_s": DT, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "hops/s", "cores": ref.cores, "kind": ref.kind, "sample": sample},
        "e2e": {"value": value, "unit": "hops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from cnt_film_monte_carlo_b200 import film
    from cnt_film_monte_carlo_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hop engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    P, n_int = args.excitons, args.intervals
    stream = torch.cuda.current_stream(dev)

    pos, ori = film.film(**film.CONFIG_FILMS["C2"])
    eng = Engine(mc_block(P), device=local_rank, stream=stream.cuda_stream)
    eng.set_mesh(pos, ori)
    for k, v in (("chunk_steps", args.chunk), ("hot_pct", args.hot_pct), ("occupancy", args.occupancy), ("top_entries", args.top_entries)):
        eng.set_option(k, v)
    for kv in args.opt:                               # any other engine option, name=value
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    if args.stage_mb > 0:
        eng.set_option("stage_mb", args.stage_mb)
    eng.kubo_init()
    eng.kubo_create_particles(P, seed=1, first_global_id=rank * P)

    sums = torch.zeros((n_int, 4), dtype=torch.float64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step():
        eng.kubo_step_dev(DT, n_int, sums.data_ptr())
        if world > 1:
            dist.all_reduce(sums)   # one NCCL all-reduce of the [intervals][4] MSD / hop histogram per step

    # warm the population up: after the first intervals excitons have left their injection sites
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    hops_total = 0.0
    msd_last = None
    for a, b in ev:
        flush.fill_(0)              # L2 flushed between timed iterations (outside the event pair)
        a.record(stream)
        one_step()
        b.record(stream)
        b.synchronize()
        hops_total += float(sums[:, 3].sum().item())   # all ranks' hops after the all-reduce
        msd_last = (sums[-1, :3] / (P * world)).tolist()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    clocks = sampler.stop() if rank == 0 else None
    value = hops_total / (ms_max * 1e-3)

    # ---- dominant kernel: duration measured live (events around every kubo_flat_kernel launch) + algorithmic bytes
    roof = None
    if rank == 0:
        eng.set_option("time_kernels", 1)
        eng.sync()
        h0 = eng.hops()
        eng.kubo_step(DT, n_int)                       # the kernel that was timed above, now with events around each launch
        k_ms, k_n = eng.kernel_ms(), eng.kernel_launches()
        hops = eng.hops() - h0
        step_ms = eng.last_step_ms()
        eng.set_option("time_kernels", 0)
        eng.set_option("stats", 1)                     # instrumented twin: counts probes and chain crossings (not timed)
        eng.sync()
        h1, c0, p0 = eng.hops(), eng.crossings(), eng.probes()
        eng.kubo_step(DT, n_int)
        hops_s = max(1, eng.hops() - h1)
        crossings = (eng.crossings() - c0) * hops / hops_s
        probes = (eng.probes() - p0) * hops / hops_s
        eng.set_option("stats", 0)
        abytes = algorithmic_bytes(hops, probes, crossings, P, k_n)
        peak, which = measured_hbm_peak()
        achieved = abytes / (k_ms * 1e-3) / 1e9
        # DRAM bytes per launch from the ncu --set full capture of this kernel on this workload
        # (profiles/round1_r44_kubo_ncu_summary.txt: 0.373 GB read + 1.286 GB written by a 36-step launch over 1e6
        # excitons = 46.1 B per exciton-step: the 32-byte (step, exciton) records plus the exciton state; the tables
        # stay in L2), scaled to this run's launch size
        traffic = NCU_DRAM_BYTES_PER_EXCITON_STEP * P * n_int / max(1, k_n)
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "ncu dram__bytes_{read,write}.sum, profiles/round1_r44_kubo_ncu_summary.txt, per exciton-step x launch size",
                "peak_source": which, "kernel": "kubo_kernel",
                "kernel_ms_per_launch": k_ms / max(1, k_n), "kernel_share_of_step": k_ms / step_ms,
                "bytes_per_hop": abytes / max(1, hops), "probes_per_hop": probes / max(1, hops),
                "crossings_per_hop": crossings / max(1, hops), "hops_per_launch": hops / max(1, k_n)}

    # ---- end to end: population in pinned host memory, uploaded and downloaded every step ---------------------------
    state = eng.particles()
    pinned = {}
    for k, v in state.items():
        t = torch.from_numpy(v).pin_memory()
        pinned[k] = t.numpy()
        pinned["_keep_" + k] = t
    msd_host = np.empty((n_int, 3))
    h2d = P * (4 + 24 + 24 + 8 + 1 + 4)
    d2h = h2d + n_int * 32
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    h_before = eng.hops()
    eng.kubo_step_host_state(DT, n_int, pinned)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    h_before = eng.hops()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.kubo_step_host_state(DT, n_int, pinned)   # synchronous: returns after the D2H
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

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = CpuReference(args.cpu_excitons, os.cpu_count() or 1)
        h20, s20 = ref.step(20)
        cpu_int = int(min(4000, max(20, 15.0 / max(s20 / 20, 1e-6))))   # about 15 s of CPU work
        hops_c, sec_c = ref.step(cpu_int)
        cpu = {"value": hops_c / sec_c, "unit": "hops/s", "cores": ref.cores, "kind": ref.kind,
               "sample": "%d excitons x %d intervals of 1e-13 s on the C2 film (%.1f s), reference built against an Armadillo stand-in" % (
                   args.cpu_excitons, cpu_int, sec_c)}

    if rank == 0:
        launches_per_step = eng.last_step_launches()
        print(json.dumps({
            "metric": "exciton hops/sec", "value": value, "unit": "hops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "excitons_per_gpu": P, "dt_s": DT, "intervals_per_step": n_int,
                       "hops_per_step": hops_total / args.steps, "l2": "flushed between timed steps (256 MiB fill, outside the events)",
                       "chunk_steps": args.chunk, "hot_pct": args.hot_pct, "occupancy": args.occupancy,
                       "parallelism": "exciton sharding x%d, tables replicated, 1 all-reduce/step" % world,
                       "msd_last_m2": msd_last},
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": "hops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--excitons", type=int, default=1_000_000, help="excitons per GPU")
    ap.add_argument("--intervals", type=int, default=100, help="sampling intervals (dt = 1e-13 s) per bench step")
    ap.add_argument("--chunk", type=int, default=64, help="time steps per kernel launch")
    ap.add_argument("--hot-pct", type=int, default=30, help="share of blocks serving the most active excitons first")
    ap.add_argument("--opt", action="append", default=[], help="extra engine option name=value (repeatable)")
    ap.add_argument("--top-entries", type=int, default=1, help="1: the three widest entries of a row are tried before the row is searched")
    ap.add_argument("--occupancy", type=int, default=5, help="resident 128-thread blocks per SM of the hop kernel (4, 5, 6, 8)")
    ap.add_argument("--stage-mb", type=int, default=0, help="cap on the (step, exciton) staging buffer in MiB (0 = engine default)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-excitons", type=int, default=8000)
    ap.add_argument("--cpu-budget", type=float, default=100.0, help="seconds of CPU work for the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
