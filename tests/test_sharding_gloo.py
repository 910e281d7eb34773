"""Multi-rank logic on the CPU (world_size 2, gloo): sharding by global exciton id gives the same trajectories as one
rank, and the all-reduced MSD equals the single-rank ensemble average.  The per-rank "engine" here is the oracle with
the engine's Philox streams; the GPU ranks do exactly the same through cntmc_kubo_step_dev + NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from cnt_film_monte_carlo_b200.parallel import shard_range
from conftest import Golden, ROOT


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, total, seed, nsteps, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from cnt_film_monte_carlo_b200.parallel import ShardedKubo
    from oracle import t1 as T1m

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = Golden("small_forster")
    sk = ShardedKubo(total, rank, world)
    t = T1m.T1()
    t.kubo_init(g.mc, g.pos_nm, g.orient)
    t.draws_philox(seed)
    t.create_particles(sk.count, first_global_id=sk.first)
    local = t.kubo_step(g.dt, nsteps) * sk.count          # un-normalised sums of this shard
    msd = sk.msd(local)                                   # all-reduce + divide by the whole population
    p = t.particles()
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), msd=msd, first=sk.first, **p)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_population():
    for total, world in ((10, 3), (1_000_000, 8), (7, 8), (100_000_000, 4)):
        spans = [shard_range(total, r, world) for r in range(world)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == total
        assert all(spans[r][0] + spans[r][1] == spans[r + 1][0] for r in range(world - 1))
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_ranks_equal_one(tmp_path):
    from oracle import t1 as T1m
    total, seed, nsteps = 61, 77, 90
    mp.spawn(worker, args=(2, free_port(), total, seed, nsteps, str(tmp_path)), nprocs=2, join=True)
    g = Golden("small_forster")
    t = T1m.T1()
    t.kubo_init(g.mc, g.pos_nm, g.orient)
    t.draws_philox(seed)
    t.create_particles(total)
    msd = t.kubo_step(g.dt, nsteps)
    whole = t.particles()
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(2)]
    for k in ("site", "pos", "delta", "ff", "heading"):
        assert np.array_equal(whole[k], np.concatenate([p[k] for p in parts], axis=-1)), k   # trajectories independent of W
    for p in parts:
        assert np.allclose(p["msd"], msd, rtol=1e-12, atol=0)                                # only the summation order moves
    assert np.array_equal(parts[0]["msd"], parts[1]["msd"])
