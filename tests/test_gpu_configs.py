"""BASELINE configs 3, 4 and 5 at their stated per-GPU sizes, through the C ABI: oracle checks on samples
and the size-independent properties the domain offers (draw accounting, shard-equals-whole, exact conservation of the
integer contact bins)."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import Engine
from cnt_film_monte_carlo_b200.parallel import ShardedContacts, shard_range
from oracle import t1 as T1m
from conftest import base_mc

pytestmark = pytest.mark.gpu
DT = 1e-13


def test_config_3_one_gpu_share_of_1e8_excitons():
    """C3: 1e8 excitons split over 8 GPUs by global id; this is rank 3's share (1.25e7 excitons, ids from 3.75e7) on the
    C2 film, stepped in full-length launches."""
    first, count = shard_range(100_000_000, 3, 8)
    assert (first, count) == (37_500_000, 12_500_000)
    pos, ori = film.film(**film.CONFIG_FILMS["C2"])
    mc = base_mc()
    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.kubo_init()
    e.kubo_create_particles(count, seed=1, first_global_id=first)
    p0 = e.particles()
    msd = e.kubo_step(DT, 40)
    p1 = e.particles()
    assert e.get_option("dbg_last_chunk") == 40                   # the whole call in one launch: no shortening by the staging budget
    assert int(p1["ndraw"].astype(np.int64).sum()) == 3 * count + 2 * e.hops() + e.reinjections()
    assert e.hops() > 0.3 * 40 * count
    off = np.abs(p1["delta"] - (p1["pos"] - p0["pos"])).max(axis=0) > 1e-18
    assert off.sum() <= e.reinjections()                         # displacement telescopes unless re-injected
    assert np.allclose(msd[-1], (p1["delta"] ** 2).mean(axis=1), rtol=1e-12, atol=0)
    # a piece of the share run alone: the same trajectories; and the oracle agrees on it
    sub_first, sub = first + 9_000_001, 2048
    s = Engine(mc)
    s.set_mesh(pos, ori)
    s.kubo_init()
    s.kubo_create_particles(sub, seed=1, first_global_id=sub_first)
    s.kubo_step(DT, 40, want_msd=False)
    ps = s.particles()
    lo = sub_first - first
    assert all(np.array_equal(ps[k], p1[k][..., lo:lo + sub]) for k in ps)
    t = T1m.T1()
    t.draws_philox(1)
    t.set_memo(True)
    t.kubo_init(mc, pos, ori)
    t.create_particles(256, first_global_id=sub_first)
    t.kubo_step(DT, 40, want_msd=False)
    pt = t.particles()
    assert np.array_equal(pt["site"], ps["site"][:256]) and np.allclose(pt["delta"], ps["delta"][:, :256], rtol=1e-9, atol=1e-18)


def test_config_4_full_size_rows_and_shard_against_oracle():
    """C4: 20 000 tubes x 250 sites = 5e6 sites, ~1.4e9 table entries (22 GB of rows in HBM).  320 sampled rows must be the
    oracle's bits (neighbour ids and cumulative rates), and a shard of the 1e6 excitons must follow the oracle's sites."""
    pos, ori = film.film(**film.CONFIG_FILMS["C4"])
    mc = base_mc()
    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.kubo_init()
    N = e.num_sites()
    assert N == 5_000_000 and e.csr_nnz() > 1.2e9 and e.csr_midpoint_guards() == 0
    t = T1m.T1()
    t.set_lazy_rates(True)      # Gamma_i when first needed: the oracle would take minutes for all 5e6 rows
    t.set_memo(True)
    t.draws_philox(1)
    t.kubo_init(mc, pos, ori)
    rng = np.random.default_rng(4)
    longest = 0
    for i in np.concatenate([[0, N - 1], rng.integers(0, N, 318)]):
        nbr, cum = e.csr_row(int(i))
        ids, c = t.row(int(i))
        assert np.array_equal(nbr, ids) and np.array_equal(cum, c), int(i)     # bit for bit
        longest = max(longest, len(nbr))
    assert longest > 400                                          # rows far beyond the 255-entry search guide
    P = 1_000_000
    e.kubo_create_particles(P, seed=1)
    e.kubo_step(DT, 64, want_msd=False)
    p = e.particles()
    assert int(p["ndraw"].astype(np.int64).sum()) == 3 * P + 2 * e.hops() + e.reinjections() and e.hops() > 0.2 * 64 * P
    first, n_o = 640_000, 96
    t.create_particles(n_o, first_global_id=first)
    t.kubo_step(DT, 64, want_msd=False)
    pt = t.particles()
    assert np.array_equal(pt["site"], p["site"][first:first + n_o])
    assert np.array_equal(pt["heading"].astype(np.uint8), p["heading"][first:first + n_o])
    assert np.allclose(pt["pos"], p["pos"][:, first:first + n_o], rtol=1e-9, atol=1e-18)


def test_config_5_one_gpu_share_of_1e9_excitons():
    """C5: contact-driven transport with ~1.25e8 excitons alive per GPU (1e9 on 8 GPUs), through parallel.ShardedContacts.
    The integer bins obey exact conservation: for every slab that touches no contact, the change of its population over
    an iteration equals the net crossings of its two interfaces; the injecting contact holds exactly its population."""
    pos, ori = film.film(**film.CONFIG_FILMS["C2"])
    mc = base_mc()
    c1_total = 8 * 22_000_000
    sc = ShardedContacts(c1_total, 0, 1, 5, 8)                    # rank 5 of 8
    e = Engine(mc)
    e.set_mesh(pos, ori)
    sc.configure(e)
    e.init(sc.c1_pop, sc.c2_pop, seed=sc.seed, capacity=int(7 * sc.c1_pop))
    assert e.number_of_particles() > 1.0e8                        # linear profile from c1_pop down to 0: 5 x c1_pop
    n_seg = e.number_of_segments()
    pop, cur = e.step(DT, 6)
    assert e.number_of_particles() > 1.0e8 and e.hops() > 1e8
    # the injecting contact was refilled to exactly c1_pop before every step; a few per cent leave its slab within one step
    assert np.all(pop[:, 0] > 0.9 * sc.c1_pop) and np.all(pop[:, 0] < 1.05 * sc.c1_pop)
    for s in range(1, 6):
        for k in range(1, n_seg - 1):
            assert pop[s, k] - pop[s - 1, k] == cur[s, k - 1] - cur[s, k], (s, k)
    # alive after the call = counted in the last iteration, minus both contact slabs, plus the fresh contact population
    assert e.number_of_particles() == pop[-1].sum() - pop[-1, 0] - pop[-1, -1] + sc.c1_pop + sc.c2_pop


def test_ten_million_site_film():
    """Mesh ingestion at scale: 40 000 tubes x 250 sites = 1e7 sites at C4's density, ~2.8e9 table
    entries (45 GB of rows).  Site records, geometry copies and row offsets are made on the device; sampled rows must be the
    oracle's bits and a short run must account for every draw."""
    import time
    cfg = dict(film.CONFIG_FILMS["C4"])
    cfg["NT"], cfg["LX"] = 2 * cfg["NT"], cfg["LX"] * 2 ** 0.5
    pos, ori = film.film(**cfg)
    mc = base_mc()
    e = Engine(mc)
    e.set_mesh(pos, ori)
    t0 = time.time()
    e.kubo_init()
    init_s = time.time() - t0
    N = e.num_sites()
    assert N == 10_000_000 and 2.4e9 < e.csr_nnz() < 4.29e9 and e.csr_midpoint_guards() == 0
    assert init_s < 20 and e.csr_build_seconds() < 2.0
    t = T1m.T1()
    t.set_lazy_rates(True)
    t.set_memo(True)
    t.draws_philox(1)
    t.kubo_init(mc, pos, ori)
    rng = np.random.default_rng(9)
    for i in np.concatenate([[0, N - 1], rng.integers(0, N, 62)]):
        nbr, cum = e.csr_row(int(i))
        ids, c = t.row(int(i))
        assert np.array_equal(nbr, ids) and np.array_equal(cum, c), int(i)
    P = 500_000
    e.kubo_create_particles(P, seed=1)
    e.kubo_step(DT, 32, want_msd=False)
    p = e.particles()
    assert int(p["ndraw"].astype(np.int64).sum()) == 3 * P + 2 * e.hops() + e.reinjections() and e.hops() > 0.2 * 32 * P
    t.create_particles(48, first_global_id=1234)
    t.kubo_step(DT, 32, want_msd=False)
    pt = t.particles()
    assert np.array_equal(pt["site"], p["site"][1234:1234 + 48])
