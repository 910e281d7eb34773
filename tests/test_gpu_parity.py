"""Parity of the CUDA engine (through the C ABI) with the golden vectors and the oracle.  Run with -m gpu on a B200."""
import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import film
from cnt_film_monte_carlo_b200.engine import CntmcError, Engine
from oracle import t1 as T1m
from conftest import base_mc

pytestmark = pytest.mark.gpu
STATE_KEYS = ("site", "pos", "delta", "ff", "heading")


def engine_for(g):
    e = Engine(g.mc)
    e.set_mesh(g.pos_nm, g.orient)
    e.kubo_init()
    return e


def test_setup_and_neighbour_table_bit_exact(golden):
    e = engine_for(golden)
    sites = e.sites()
    for k, v in golden.group("site_").items():
        assert np.array_equal(sites[k], v), k
    assert np.array_equal(e.domain(), golden.z["domain"]) and np.array_equal(e.removal_domain(), golden.z["removal"])
    assert np.array_equal(e.inject(), golden.z["inject"])
    tab = e.rate_table()
    for k, v in golden.group("table_").items():
        assert np.array_equal(tab[k], v), k
    rp, nbr, cum = e.csr()
    assert np.array_equal(rp, golden.z["row_ptr"])
    assert np.array_equal(nbr, golden.z["nbr"])       # neighbour lists: same members, same order
    assert np.array_equal(cum, golden.z["cum"])       # cumulative rates: bit for bit (<= 1e-12 rel required)
    assert e.csr_midpoint_guards() == 0


def test_replay_of_reference_draws_is_bit_exact(golden):
    e = engine_for(golden)
    e.kubo_create_particles_replay(golden.z["draw_off"], golden.z["draws"], golden.z["draw_logs"])
    p0 = e.particles()
    for k in ("site", "pos", "ff", "heading"):
        assert np.array_equal(p0[k], golden.z["p0_" + k]), k
    msd = e.kubo_step(golden.dt, golden.nsteps)
    p1 = e.particles()
    for k in STATE_KEYS:
        assert np.array_equal(p1[k], golden.z["p1_" + k]), k
    assert np.array_equal(p1["ndraw"], np.diff(golden.z["draw_off"]))
    assert np.allclose(msd, golden.z["msd"], rtol=1e-12, atol=0)
    assert e.time() == float(golden.z["time"])


def test_replay_with_device_log_keeps_site_sequences(golden):
    """Without the host's log values the free-flight times may differ in the last place; sites must not."""
    e = engine_for(golden)
    e.kubo_create_particles_replay(golden.z["draw_off"], golden.z["draws"], None)
    e.kubo_step(golden.dt, golden.nsteps, want_msd=False)
    p1 = e.particles()
    assert np.array_equal(p1["site"], golden.z["p1_site"]) and np.array_equal(p1["heading"], golden.z["p1_heading"])
    assert np.allclose(p1["pos"], golden.z["p1_pos"], rtol=1e-9, atol=1e-18)
    assert np.allclose(p1["ff"], golden.z["p1_ff"], rtol=1e-9, atol=1e-24)


def test_replay_list_too_short_is_reported(golden_small):
    g = golden_small
    e = engine_for(g)
    off = g.z["draw_off"].copy()
    cut = off.copy()
    cut[1:] = np.minimum(off[1:], off[:-1] + 4)  # at most 4 draws per exciton
    draws = np.concatenate([g.z["draws"][off[i]:cut[i + 1]] for i in range(g.P)])
    noff = np.zeros_like(off)
    np.cumsum(cut[1:] - off[:-1], out=noff[1:])
    e.kubo_create_particles_replay(noff, draws, None)
    with pytest.raises(CntmcError) as ei:
        e.kubo_step(g.dt, g.nsteps)
    assert ei.value.code == -4


def test_philox_matches_oracle_site_sequences_and_state(golden):
    P, nsteps = 96, 120
    e = engine_for(golden)
    e.kubo_create_particles(P, seed=777, first_global_id=5000)
    e.trace_enable(1 << 13)
    msd = e.kubo_step(golden.dt, nsteps)
    t = T1m.T1()
    t.kubo_init(golden.mc, golden.pos_nm, golden.orient)
    t.draws_philox(777)
    t.trace_sites(True)
    t.create_particles(P, first_global_id=5000)
    msd_t = t.kubo_step(golden.dt, nsteps)
    pe, pt = e.particles(), t.particles()
    assert np.array_equal(pe["site"], pt["site"]) and np.array_equal(pe["heading"], pt["heading"].astype(np.uint8))
    # device log() vs glibc log() may differ in the last place of a free-flight time
    assert np.allclose(pe["pos"], pt["pos"], rtol=1e-9, atol=1e-18) and np.allclose(pe["ff"], pt["ff"], rtol=1e-9, atol=1e-24)
    counts, sites = e.trace()
    off_t, flat_t = t.traced_sites(5000 + P)
    assert np.array_equal(counts, np.diff(off_t)[5000:])
    flat_e = np.concatenate([sites[i, :counts[i]] for i in range(P)])
    assert np.array_equal(flat_e, flat_t)          # every scattering event lands on the same site, in the same order
    assert e.hops() == t.hops() and e.reinjections() == t.reinjections()
    assert np.allclose(msd, msd_t, rtol=1e-9, atol=0)


def test_chunking_scheduling_and_occupancy_do_not_change_results():
    pos, ori = film.film(NT=150, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
    mc = base_mc()
    states, msds = [], []
    # the first run is the plain loop (no trap solver); every other scheduling must give the same bits
    for opts in (dict(chunk_steps=64, occupancy=6, deep_thr=0), dict(chunk_steps=7, occupancy=4, deep_thr=16), dict(chunk_steps=1, deep_thr=8, deep_rounds=1),
                 dict(chunk_steps=200, occupancy=5, deep_thr=0), dict(chunk_steps=31, occupancy=7, deep_thr=0), dict(chunk_steps=50, occupancy=8, deep_thr=0), dict(chunk_steps=64, occupancy=0, hot_pct=10, deep_thr=0), dict(chunk_steps=64, stage_mb=1, deep_thr=16, deep_blocks=1),
                 dict(chunk_steps=8, hot_pct=0, deep_thr=16), dict(chunk_steps=3, hot_pct=100, deep_thr=16, deep_blocks=1),
                 dict(chunk_steps=8, hot_pct=30, occupancy=6, deep_thr=1, deep_blocks=2),
                 dict(top_entries=0), dict(top_entries=1, chunk_steps=16, deep_thr=0), dict(top_entries=0, runs=0, hot_pct=50),
                 dict(dirs=0, chunk_steps=9, deep_thr=30), dict(chunk_steps=40, deep_thr=2, deep_rounds=2), dict(chunk_steps=64, deep_thr=9, deep_rounds=1),
                 dict(chunk_steps=64, deep_thr=8, deep_group=1), dict(chunk_steps=11, deep_thr=16, deep_group=1, deep_rounds=1), dict(runs=0, hot_pct=5, chunk_steps=33, deep_thr=16),
                 # event bursts; the trap kernel beside the lane kernel of the same launch (in-launch hand-over)
                 dict(chunk_steps=64, deep_thr=8, deep_group=1, trap_burst=4), dict(chunk_steps=64, deep_thr=8, deep_group=1, trap_burst=4, deep_overlap=1),
                 dict(chunk_steps=9, deep_thr=16, deep_group=1, deep_overlap=1, overlap_trap_blocks=2, trap_burst=3),
                 dict(chunk_steps=64, deep_thr=1, deep_group=1, deep_overlap=1, trap_burst=64, hot_pct=50)):
        e = Engine(mc)
        e.set_mesh(pos, ori)
        for k, v in opts.items():
            e.set_option(k, v)
        e.kubo_init()
        e.kubo_create_particles(5000, seed=4)
        m = [e.kubo_step(1e-13, 50), e.kubo_step(1e-13, 37)]
        states.append(e.particles())
        msds.append(np.concatenate(m))
        hops = e.hops()
    for s in states[1:]:
        assert all(np.array_equal(s[k], states[0][k]) for k in s)  # per-exciton results: bit-identical
    for m in msds[1:]:
        assert np.array_equal(m, msds[0])                          # ensemble sums: fixed-order reduction over exciton index
    assert hops > 5000


def test_walk_square_root_and_divisions_are_the_ieee_operations():
    """The chain walk's call-free square root and divisions (csrc/hop_core.h sqrt_walk / div3_walk) against sqrt() and '/' on the
    device: 1e9 operand sets over the accepted exponent range (perfect squares, zeros of both signs, |w| = d included) -- no bit
    differs, nothing in range is flagged, everything outside the range is."""
    import ctypes
    from cnt_film_monte_carlo_b200 import _lib
    counts = (ctypes.c_int64 * 4)()
    rc = _lib.load().cntmc_dbg_walk_arith(0, 1_000_000_000, 20261018, counts)
    assert rc == 0, _lib.load().cntmc_last_error(None)
    assert list(counts) == [0, 0, 0, 0]


def test_shortcuts_do_not_change_results_on_a_trimmed_film():
    """Top entries, segment-time chain walks and stored directions against the plain paths, and all of them against the
    oracle, on a film whose tubes are cut by the trim box (permuted site order, broken runs, reflections at chain ends)."""
    pos, ori = film.film(NT=120, NP=80, a=5.0, LX=300.0, LY=60.0, seed=3)
    mc = base_mc(**{"trim limits": {"xlim": [2e-8, 2.6e-7], "ylim": [0.0, 5e-8], "zlim": [1e-8, 2.8e-7]}})
    res = []
    for opts in (dict(top_entries=0, runs=0, dirs=0, deep_thr=0), dict(top_entries=1, runs=1, dirs=1, deep_thr=16), dict(top_entries=1, runs=1, dirs=0, deep_thr=0),
                 dict(top_entries=0, runs=0, dirs=1, deep_thr=8, deep_blocks=3), dict(top_entries=1, runs=1, dirs=1, chunk_steps=5, deep_thr=12, deep_rounds=1)):
        e = Engine(mc)
        e.set_mesh(pos, ori)
        for k, v in opts.items():
            e.set_option(k, v)
        e.kubo_init()
        e.kubo_create_particles(3000, seed=11)
        e.trace_enable(1 << 11)
        msd = np.concatenate([e.kubo_step(1e-13, 70), e.kubo_step(7e-13, 30)])
        counts, sites = e.trace()
        assert counts.max() < (1 << 11)
        flat = np.concatenate([sites[i, :counts[i]] for i in range(len(counts))])
        res.append((e.particles(), msd, counts, flat, e.hops(), e.reinjections()))
    for r in res[1:]:
        assert all(np.array_equal(r[0][k], res[0][0][k]) for k in r[0])
        assert np.array_equal(r[1], res[0][1]) and np.array_equal(r[2], res[0][2]) and r[4:] == res[0][4:]
        assert np.array_equal(r[3], res[0][3])
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    t.draws_philox(11)
    t.create_particles(3000)
    t.kubo_step(1e-13, 70, want_msd=False)
    t.kubo_step(7e-13, 30, want_msd=False)
    pt, pe = t.particles(), res[1][0]
    assert np.array_equal(pe["site"], pt["site"]) and np.array_equal(pe["heading"], pt["heading"].astype(np.uint8))
    assert np.allclose(pe["pos"], pt["pos"], rtol=1e-9, atol=1e-18) and res[1][4] == t.hops() and res[1][5] == t.reinjections()


def test_host_state_call_equals_resident_call(golden_small):
    g = golden_small
    a, b = engine_for(g), engine_for(g)
    a.kubo_create_particles(300, seed=21)
    b.kubo_create_particles(300, seed=21)
    state = b.particles()
    ma = a.kubo_step(g.dt, 40)
    mb = b.kubo_step_host_state(g.dt, 40, state)
    pa = a.particles()
    assert all(np.array_equal(pa[k], state[k]) for k in pa)
    assert np.allclose(ma, mb, rtol=1e-13, atol=0)


def test_population_split_over_handles_gives_same_trajectories(golden_small):
    """Exciton streams are keyed by global id: two half-populations == one whole (the multi-GPU sharding rule)."""
    g = golden_small
    whole = engine_for(g)
    whole.kubo_create_particles(200, seed=8, first_global_id=0)
    whole.kubo_step(g.dt, 80, want_msd=False)
    pw = whole.particles()
    parts = []
    for first in (0, 100):
        e = engine_for(g)
        e.kubo_create_particles(100, seed=8, first_global_id=first)
        e.kubo_step(g.dt, 80, want_msd=False)
        parts.append(e.particles())
    for k in pw:
        assert np.array_equal(pw[k], np.concatenate([p[k] for p in parts], axis=-1)), k


def test_c2_size_table_against_oracle():
    """BASELINE config 2's film (1000 tubes x 100 sites): 3.08e6 CSR entries, all bit-identical to the oracle."""
    pos, ori = film.film(**film.CONFIG_FILMS["C2"])
    mc = base_mc()
    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.kubo_init()
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    for a, b in zip(e.csr(), t.csr()):
        assert np.array_equal(a, b)
    se, st = e.sites(), t.sites()
    assert np.array_equal(se["max_rate"], st["max_rate"]) and np.array_equal(se["inv_max_rate"], st["inv_max_rate"])
    assert e.csr_midpoint_guards() == 0
    # 20 000 excitons, 30 steps: site of every exciton identical to the oracle (memoised rows)
    t.set_memo(True)
    t.draws_philox(1)
    t.create_particles(20000)
    t.kubo_step(1e-13, 30, want_msd=False)
    e.kubo_create_particles(20000, seed=1)
    e.kubo_step(1e-13, 30, want_msd=False)
    pe, pt = e.particles(), t.particles()
    assert np.array_equal(pe["site"], pt["site"])
    assert np.allclose(pe["delta"], pt["delta"], rtol=1e-9, atol=1e-18)
    assert e.hops() == t.hops()


def test_isolated_site_is_an_error_not_garbage():
    pos, ori = film.film(NT=3, NP=1, a=5.0, LX=500.0, LY=300.0, seed=2)  # three far-apart single-site "tubes"
    e = Engine(base_mc())
    e.set_mesh(pos, ori)
    with pytest.raises(CntmcError) as ei:
        e.kubo_init()
    assert ei.value.code == -3 and "no neighbour" in str(ei.value)


def test_dense_film_long_rows_against_oracle():
    """BASELINE config 4's density at 1/50 of its size: 2 nm segmentation, rows of several hundred entries (beyond the
    255-entry search guide), L2-spilling table.  Sampled rows and a short run must equal the oracle."""
    pos, ori = film.film(NT=400, NP=250, a=2.0, LX=283.0, LY=100.0, seed=1234)
    mc = base_mc()
    e = Engine(mc)
    e.set_mesh(pos, ori)
    e.kubo_init()
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    rp, nbr, cum = e.csr()
    deg = np.diff(rp)
    assert deg.mean() > 150 and deg.max() > 255
    rng = np.random.default_rng(0)
    rows = np.concatenate([rng.integers(0, len(deg), 300), np.argsort(deg)[-20:], np.argsort(deg)[:20]])
    for i in rows:
        ids, c = t.row(int(i))
        assert np.array_equal(ids, nbr[rp[i]:rp[i + 1]]) and np.array_equal(c, cum[rp[i]:rp[i + 1]]), i
    st = t.sites()
    assert np.array_equal(e.sites()["max_rate"], st["max_rate"])
    t.set_memo(True)
    t.draws_philox(2)
    t.create_particles(3000)
    t.kubo_step(1e-13, 12, want_msd=False)
    e.kubo_create_particles(3000, seed=2)
    e.kubo_step(1e-13, 12, want_msd=False)
    assert np.array_equal(e.particles()["site"], t.particles()["site"]) and e.hops() == t.hops()


def test_host_state_refuses_what_would_read_out_of_bounds(golden_small):
    """cntmc_kubo_step_host_state indexes the site tables with the caller's arrays: a bad site or a non-finite free-flight
    time is an error, not a device fault; and the handle stays usable."""
    g = golden_small
    e = engine_for(g)
    e.kubo_create_particles(64, seed=2)
    state = e.particles()
    bad = {k: v.copy() for k, v in state.items()}
    bad["site"][5] = e.num_sites()
    with pytest.raises(CntmcError, match="site index out of range"):
        e.kubo_step_host_state(g.dt, 3, bad)
    bad = {k: v.copy() for k, v in state.items()}
    bad["ff"][7] = np.nan
    with pytest.raises(CntmcError, match="not finite"):
        e.kubo_step_host_state(g.dt, 3, bad)
    e.kubo_step_host_state(g.dt, 3, state)


def test_trace_capacity_is_a_clamp_not_a_buffer_size(golden_small):
    """An exciton with more events than the trace holds is counted, not recorded: nothing is written past its slots,
    neither within a launch nor when the next launch starts beyond the capacity."""
    g = golden_small
    full, tiny = engine_for(g), engine_for(g)
    for e, cap in ((full, 1 << 12), (tiny, 4)):
        e.kubo_create_particles(200, seed=13)
        e.trace_enable(cap)
        e.kubo_step(g.dt, 90, want_msd=False)
        e.kubo_step(g.dt, 60, want_msd=False)
    cf, sf = full.trace()
    ct, st = tiny.trace()
    assert np.array_equal(cf, ct) and cf.max() > 4          # counts run on beyond the capacity
    assert np.array_equal(st, sf[:, :4])                    # the first `cap` events of every exciton, nothing else
    pf, pt = full.particles(), tiny.particles()
    assert all(np.array_equal(pf[k], pt[k]) for k in pf)


def test_a_new_population_can_follow_a_failed_replay(golden_small):
    """Device error flags are reported once: a replay list that runs out fails the call (CNTMC_ERR_REPLAY), and a fresh
    population on the same handle runs without re-initialising it."""
    g = golden_small
    e = engine_for(g)
    off, draws = g.z["draw_off"].astype(np.int64), g.z["draws"]
    cut = off.copy()
    cut[1:] = off[:-1] + np.minimum(np.diff(off), 4)      # at most four draws per exciton: creation passes, stepping runs dry
    flat = np.concatenate([draws[off[i]:cut[i + 1]] for i in range(len(off) - 1)])
    noff = np.zeros_like(off)
    noff[1:] = np.cumsum(cut[1:] - off[:-1])
    e.kubo_create_particles_replay(noff, flat, None)
    with pytest.raises(CntmcError) as ei:
        e.kubo_step(g.dt, int(g.nsteps))
    assert ei.value.code == -4
    e.kubo_create_particles(50, seed=3)
    e.kubo_step(g.dt, 20)
    assert e.hops() > 0


def test_rows_next_to_a_theta_midpoint_are_recomputed_with_glibc(golden):
    """The one place where the device's libm could change a table index relative to the reference is an acos that lands on
    a grid midpoint.  Such rows are recomputed on the host with glibc's acos and patched in.  None occurs on any film
    here, so the band is widened (option guard_ppb: 0.03 grid pitches instead of 1e-9) to exercise the repair: hundreds of
    rows go through the host path, and the table is still the reference's, bit for bit."""
    e = Engine(golden.mc)
    e.set_mesh(golden.pos_nm, golden.orient)
    e.set_option("guard_ppb", 30_000_000)
    e.kubo_init()
    assert e.csr_midpoint_guards() > 50 and e.get_option("dbg_midpoint_repairs") == e.csr_midpoint_guards()
    assert e.get_option("dbg_midpoint_changed") == 0              # the device's acos picked the same indices as glibc's
    row_ptr, nbr, cum = e.csr()
    assert np.array_equal(row_ptr, golden.z["row_ptr"]) and np.array_equal(nbr, golden.z["nbr"]) and np.array_equal(cum, golden.z["cum"])
    plain = Engine(golden.mc)
    plain.set_mesh(golden.pos_nm, golden.orient)
    plain.kubo_init()
    assert plain.csr_midpoint_guards() == 0


def test_table_build_one_warp_per_row_equals_one_thread_per_row():
    """The fill pass of the table build, one warp per row (default) against the one-thread-per-row kernel: every entry, every
    site's rate fields, search guide and top entries (seen through a run with and without the shortcuts that read them)."""
    pos, ori = film.film(NT=300, NP=250, a=2.0, LX=245.0, LY=100.0, seed=1234)   # C4's density: rows of several hundred entries
    mc = base_mc()
    out = []
    for warp in (1, 0):
        e = Engine(mc)
        e.set_mesh(pos, ori)
        e.set_option("csr_warp", warp)
        e.kubo_init()
        rp, nbr, cum = e.csr()
        s = e.sites()
        e.kubo_create_particles(4000, seed=6)
        e.trace_enable(1 << 12)
        msd = e.kubo_step(1e-13, 40)
        counts, sites = e.trace()
        out.append((rp, nbr, cum, s["max_rate"], s["inv_max_rate"], msd, counts, sites, e.particles(), e.get_option("dbg_top_events")))
    a, b = out
    assert np.diff(a[0]).max() > 400
    for k in range(8):
        assert np.array_equal(a[k], b[k]), k
    assert all(np.array_equal(a[8][k], b[8][k]) for k in a[8]) and a[9] == b[9] and a[9] > 0
