"""Sharding of the exciton population over the GPUs of one box.

Excitons never interact and all tables are read-only while stepping, so rank r of W simply owns the excitons with
global ids [first, first + count): their counter-based streams are keyed by global id, hence every trajectory is
the same for any W.  The only exchange is one all-reduce (sum) of the per-interval [sum dx^2, sum dy^2, sum dz^2, hops]
rows (Green-Kubo) or of the integer population / current bins (contacts) per engine call -- a few kilobytes over
NCCL/NVLink.  torch.distributed is the plumbing; nothing here computes.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global id, count) of rank's share of `total` excitons; the first total % world ranks hold one more."""
    base, extra = divmod(int(total), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def allreduce_sums(local, group=None):
    """Sum a [nsteps][k] array of un-normalised sums over all ranks.  numpy in, numpy out (gloo, CPU tests); a CUDA
    torch tensor is reduced in place over NCCL and returned."""
    import torch
    import torch.distributed as dist

    if isinstance(local, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(local))
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(t, group=group)
        return t.numpy()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local, group=group)
    return local


class ShardedKubo:
    """One rank's part of a Green-Kubo run: `stepper(dt, nsteps)` must return this rank's [nsteps][3] sums of squared
    displacements (e.g. ``lambda dt, n: engine.kubo_step(dt, n) * count``)."""

    def __init__(self, total_excitons: int, rank: int, world: int, group=None):
        self.total, self.rank, self.world, self.group = int(total_excitons), rank, world, group
        self.first, self.count = shard_range(total_excitons, rank, world)

    def msd(self, local_sums: np.ndarray) -> np.ndarray:
        """Ensemble average over the WHOLE population from this rank's local sums (monte_carlo.cpp:402-404)."""
        return allreduce_sums(np.asarray(local_sums, np.float64), self.group) / float(self.total)


class ShardedContacts:
    """One rank's part of a contact-mode run (monte_carlo::init / step / save_metrics / repopulate_contacts).

    The contact rules are per exciton, so the populations the contacts are held at are split statically over the ranks
    (`c1_pop`, `c2_pop` = this rank's share, the first ranks hold the remainders) and every rank runs its own excitons
    on a full copy of the film.  Streams must differ between ranks: `seed` mixes the rank into the run's seed, and
    `configure(engine)` moves the rank's exciton ids to their own range (rank * 2^56 onwards), which alone guarantees that
    no two ranks ever share a stream (the stream key is a bijection of the id's high half for a fixed seed).  Unlike
    the Green-Kubo flavour the trajectories depend on the number of ranks (birth order defines the exciton ids); the
    ensemble is the same.  The one exchange per engine call is the sum of the integer bins
    [nsteps][n_seg populations + (n_seg - 1) net crossings] (monte_carlo.h:566-573, 626-636)."""

    def __init__(self, c1_pop: int, c2_pop: int, seed: int, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.c1_pop = shard_range(c1_pop, rank, world)[1]
        self.c2_pop = shard_range(c2_pop, rank, world)[1]
        self.seed = (int(seed) + 0x9E3779B97F4A7C15 * rank) & 0xFFFFFFFFFFFFFFFF

    def configure(self, engine) -> None:
        """Call before engine.init(): this rank's excitons are numbered from rank * 2^56."""
        engine.set_option("gid_base_shift56", self.rank)

    def bins(self, local_bins):
        """Whole-ensemble bins from this rank's (numpy [nsteps][2*n_seg-1] int64, or a CUDA tensor filled by
        cntmc_step_dev, reduced in place over NCCL)."""
        if isinstance(local_bins, np.ndarray):
            return allreduce_sums(np.asarray(local_bins, np.int64), self.group)
        return allreduce_sums(local_bins, self.group)
