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


def contact_worker(rank, world, port, c1, c2, seed, nsteps, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from cnt_film_monte_carlo_b200.parallel import ShardedContacts
    from oracle import t1 as T1m

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = Golden("small_forster")
    sc = ShardedContacts(c1, c2, seed, rank, world)
    t = T1m.T1()
    t.draws_philox(sc.seed)
    t.set_memo(True)
    t.contacts_init(g.mc, g.pos_nm, g.orient, c1_pop=sc.c1_pop, c2_pop=sc.c2_pop)
    alive = [t.L.t1_num_particles(t.h)]
    local = []
    for _ in range(nsteps):
        pop, cur = t.contact_iteration(1e-14)
        local.append(np.concatenate([pop, cur]))
        alive.append(t.L.t1_num_particles(t.h))
    local = np.array(local, np.int64)
    total = sc.bins(local.copy())
    np.savez(os.path.join(out_dir, f"contact_rank{rank}.npz"), local=local, total=total, alive=np.array(alive), c1=sc.c1_pop, c2=sc.c2_pop,
             seed=np.uint64(sc.seed))
    dist.barrier()
    dist.destroy_process_group()


def test_contact_mode_two_ranks(tmp_path):
    """Contact populations are split over the ranks, the integer bins are all-reduced: every exciton alive on any rank is
    counted exactly once per step, both ranks see the same totals, and the shares add up to the requested populations."""
    c1, c2, nsteps = 301, 40, 30
    mp.spawn(contact_worker, args=(2, free_port(), c1, c2, 5, nsteps, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"contact_rank{r}.npz") for r in range(2)]
    g = Golden("small_forster")
    n_seg = int(g.mc["number of segments"])
    assert int(parts[0]["c1"]) + int(parts[1]["c1"]) == c1 and int(parts[0]["c2"]) + int(parts[1]["c2"]) == c2
    assert int(parts[0]["seed"]) != int(parts[1]["seed"])
    assert np.array_equal(parts[0]["total"], parts[1]["total"])
    assert np.array_equal(parts[0]["total"], parts[0]["local"] + parts[1]["local"])
    alive = parts[0]["alive"] + parts[1]["alive"]
    assert np.array_equal(parts[0]["total"][:, :n_seg].sum(axis=1), alive[:-1])      # counted before the contacts are refilled
    assert not np.array_equal(parts[0]["local"], parts[1]["local"])                   # different streams
