"""The engine's host/device-shared arithmetic (csrc/hop_core.h, csr_core.h) and host set-up (host_setup.cpp), compiled
for the CPU by tests/host_emul, against the golden vectors and the oracle.  This is the code the CUDA kernels inline;
the GPU tests (-m gpu) repeat these checks through the real kernels."""
import numpy as np
import pytest

from emul import Emul
from oracle import t1 as T1m
from cnt_film_monte_carlo_b200 import film
from conftest import base_mc

STATE_KEYS = ("site", "pos", "delta", "ff", "heading")


def test_setup_and_neighbour_table(golden):
    e = Emul(golden.mc)
    e.kubo_init(golden.pos_nm, golden.orient)
    sites = e.sites()
    for k, v in golden.group("site_").items():
        assert np.array_equal(sites[k], v), k  # includes the closed-form trim permutation and Gamma_i, 1/Gamma_i
    dom, rem = e.domains()
    assert np.array_equal(dom, golden.z["domain"]) and np.array_equal(rem, golden.z["removal"])
    assert np.array_equal(e.inject(), golden.z["inject"])
    rp, nbr, cum = e.csr()
    assert np.array_equal(rp, golden.z["row_ptr"]) and np.array_equal(nbr, golden.z["nbr"])
    assert np.array_equal(cum, golden.z["cum"])
    assert np.array_equal(e.table_rates(golden.z["table_rates"].size), golden.z["table_rates"].ravel())
    assert e.guards() == 0


def test_replay_is_bit_exact(golden):
    e = Emul(golden.mc)
    e.kubo_init(golden.pos_nm, golden.orient)
    e.create_replay(golden.z["draw_off"], golden.z["draws"], golden.z["draw_logs"])
    p0 = e.particles()
    for k in ("site", "pos", "ff", "heading"):
        assert np.array_equal(p0[k], golden.z["p0_" + k]), k
    msd = e.kubo_step(golden.dt, golden.nsteps)
    p1 = e.particles()
    for k in STATE_KEYS:
        assert np.array_equal(p1[k], golden.z["p1_" + k]), k
    assert np.array_equal(p1["ndraw"], np.diff(golden.z["draw_off"]))  # every recorded draw consumed, none missing
    assert np.allclose(msd, golden.z["msd"], rtol=1e-13, atol=0)


def test_step_subdivision_is_exact(golden_small):
    """Running the steps one call at a time changes nothing (state carries everything)."""
    g = golden_small
    a, b = Emul(g.mc), Emul(g.mc)
    for e in (a, b):
        e.kubo_init(g.pos_nm, g.orient)
        e.create_philox(40, seed=9)
    a.kubo_step(g.dt, 60)
    for _ in range(60):
        b.kubo_step(g.dt, 1)
    pa, pb = a.particles(), b.particles()
    assert all(np.array_equal(pa[k], pb[k]) for k in pa)


def test_philox_matches_oracle_with_site_sequences(golden):
    e = Emul(golden.mc)
    e.kubo_init(golden.pos_nm, golden.orient)
    e.create_philox(64, seed=12345, first_gid=1000)
    msd = e.kubo_step(golden.dt, 150, trace_cap=1 << 14)
    t = T1m.T1()
    t.kubo_init(golden.mc, golden.pos_nm, golden.orient)
    t.draws_philox(12345)
    t.trace_sites(True)
    t.create_particles(64, first_global_id=1000)
    msd_t = t.kubo_step(golden.dt, 150)
    pe, pt = e.particles(), t.particles()
    for k in STATE_KEYS:
        assert np.array_equal(pe[k], pt[k]), k
    off_t, flat_t = t.traced_sites(1064)
    off_e, flat_e = e.trace()
    assert np.array_equal(np.diff(off_t)[1000:], np.diff(off_e)) and np.array_equal(flat_t, flat_e)
    assert e.hops() == t.hops() and len(flat_e) == e.hops()
    assert np.allclose(msd, msd_t, rtol=1e-13, atol=0)


def test_select_entry_equals_reference_loop():
    import ctypes
    e = Emul(base_mc())
    e.L.emul_select_full.restype = ctypes.c_int64
    e.L.emul_select_full.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double]
    rng = np.random.default_rng(11)
    for _ in range(3000):
        d = int(rng.integers(1, 120))
        c = np.cumsum(rng.random(d) * (rng.random(d) > 0.25))
        for dice in (0.0, c[-1], c[-1] * rng.random(), float(rng.choice(c)), np.nextafter(float(rng.choice(c)), 0)):
            assert e.select(c, dice) == T1m.select(c, dice)
            assert e.L.emul_select_full(c.ctypes.data, d, ctypes.c_double(dice)) == T1m.select(c, dice)  # the engine's entry search


def test_top_entries_decide_exactly_what_the_reference_loop_decides():
    """The shortcut of the event path: a draw inside the stored draw interval of one of the three widest entries selects that
    entry -- exactly the entry the reference's comparison of doubles selects for the dice of that draw.  (Intervals are stored in
    whole blocks of 2^16 draws, rounded inwards, so draws next to an interval's end are left to the ordinary search.)"""
    import ctypes
    e = Emul(base_mc())
    e.L.emul_select_top.restype = ctypes.c_int64
    e.L.emul_select_top.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)]
    rng = np.random.default_rng(23)
    decided = undecided = 0
    RAND_MAX = 2147483647
    for _ in range(3000):
        d = int(rng.integers(1, 60))
        rates = rng.random(d) * (rng.random(d) > 0.25)
        rates[rng.integers(0, d, size=int(rng.integers(0, 4)))] *= 1e3   # a few dominant entries, as on hot sites
        c = np.cumsum(rates)
        if c[-1] == 0.0:
            continue
        # draws at and around the images of the interval ends (block boundaries included), the ends of the range, random ones
        rs = {0, 1, 65535, 65536, RAND_MAX - 65536, RAND_MAX - 1, RAND_MAX} | {int(x) for x in rng.integers(0, RAND_MAX + 1, size=4)}
        for x in rng.choice(c, size=3):
            r0 = int(min(RAND_MAX, max(0, round(float(x) / c[-1] * RAND_MAX))))
            for r in (r0 - 65536, r0 - 1, r0, r0 + 1, r0 + 65536, (r0 >> 16) << 16, ((r0 >> 16) << 16) - 1, ((r0 >> 16) + 1) << 16):
                if 0 <= r <= RAND_MAX:
                    rs.add(r)
        for r in rs:
            dice = ctypes.c_double()
            k = e.L.emul_select_top(c.ctypes.data, d, r, ctypes.byref(dice))
            if k >= 0:
                assert k == T1m.select(c, dice.value)
                decided += 1
            else:
                undecided += 1
    assert decided > 10000 and undecided > 1000


def test_draw_block_intervals_are_rounded_inwards_and_as_wide_as_possible():
    """TopRec stores [rlo, rhi) as whole blocks of 2^16 draws: every draw the packed interval accepts lies in [rlo, rhi), and the
    blocks just outside it do not fit (so only draws in the two boundary blocks are left to the ordinary search)."""
    import ctypes
    e = Emul(base_mc())
    pack, member = e.L.emul_draw_blocks, e.L.emul_in_draw_blocks
    pack.restype = ctypes.c_uint32
    pack.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    member.restype = ctypes.c_int
    member.argtypes = [ctypes.c_uint32, ctypes.c_uint32]
    rng = np.random.default_rng(99)
    B = 1 << 16
    cases = [(0, 0), (0, 1 << 31), (0, B), (1, B), (0, B - 1), (B, 2 * B), (B - 1, 2 * B + 1), ((1 << 31) - B, 1 << 31), ((1 << 31) - 1, 1 << 31), (1 << 31, 1 << 31)]
    for _ in range(20000):
        lo = int(rng.integers(0, (1 << 31) + 1))
        hi = int(min(1 << 31, lo + int(rng.integers(0, 1 << int(rng.integers(1, 32))))))
        cases.append((lo, hi))
    for lo, hi in cases:
        iv = pack(lo, hi)
        blo, n = iv & 0xffff, iv >> 16
        assert blo * B >= lo and (n == 0 or (blo + n) * B <= hi)                  # inside [rlo, rhi)
        if n:
            assert (blo - 1) * B < lo or blo == 0 and lo == 0 or (blo - 1) * B < lo  # the block before does not fit
            assert (blo + n + 1) * B > hi                                            # nor the block after
        else:
            assert hi - lo < 2 * B - 1 or hi <= lo                                   # no whole block fits
        for r in {lo, max(lo, hi - 1), (lo + hi) // 2, blo * B, max(0, blo * B - 1), min((1 << 31) - 1, (blo + n) * B), (blo + n) * B - 1}:
            if 0 <= r < (1 << 31):
                if member(iv, r):
                    assert lo <= r < hi
                elif n:
                    assert not (blo * B <= r < (blo + n) * B)


def test_first_draw_reaching_is_the_smallest_draw_whose_dice_reaches_the_bound():
    import ctypes
    e = Emul(base_mc())
    f = e.L.emul_first_draw_reaching
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_double, ctypes.c_double]
    rng = np.random.default_rng(5)
    assert f(1.0, -1.0) == 0 and f(1.0, 0.0) == 0 and f(3.7e12, 3.8e12) == 1 << 31
    for _ in range(20000):
        total = float(10.0 ** rng.uniform(-3, 15) * rng.uniform(1, 10))
        x = total * float(rng.random()) if rng.random() < 0.9 else float(np.nextafter(total, rng.choice([0.0, np.inf])))
        assert f(total, x) >= 0


def test_top_entries_path_is_bit_identical_and_matches_oracle(golden):
    """Events decided by the top entries of the row change nothing."""
    res = []
    for on in (False, True):
        e = Emul(golden.mc)
        e.kubo_init(golden.pos_nm, golden.orient)
        e.set_top_entries(on)
        e.create_philox(96, seed=77, first_gid=500)
        msd = e.kubo_step(golden.dt, 150, trace_cap=1 << 14)
        res.append((e.particles(), msd, e.trace(), e.hops(), e.top_events()))
    (pa, ma, (oa, fa), ha, _), (pb, mb, (ob, fb), hb, nfast) = res
    assert all(np.array_equal(pa[k], pb[k]) for k in pa) and np.array_equal(ma, mb)
    assert np.array_equal(oa, ob) and np.array_equal(fa, fb) and ha == hb
    assert nfast > 0.3 * hb, (nfast, hb)   # the shortcut is actually taken
    t = T1m.T1()
    t.kubo_init(golden.mc, golden.pos_nm, golden.orient)
    t.draws_philox(77)
    t.trace_sites(True)
    t.create_particles(96, first_global_id=500)
    t.kubo_step(golden.dt, 150)
    pt = t.particles()
    for k in STATE_KEYS:
        assert np.array_equal(pb[k], pt[k]), k
    off_t, flat_t = t.traced_sites(596)
    assert np.array_equal(np.diff(off_t)[500:], np.diff(ob)) and np.array_equal(flat_t, fb)


def test_chain_walks_on_segment_times_are_bit_identical(golden):
    """fly() over runs of memory-consecutive sites (Tables::seg) against the record-by-record walk: a film whose tubes
    are cut by the trim box (permuted site order, broken runs, reflections at chain ends) and the golden film."""
    pos, ori = film.film(NT=120, NP=80, a=5.0, LX=300.0, LY=60.0, seed=3)
    mc = dict(base_mc())
    mc["trim limits"] = {"xlim": [2e-8, 2.6e-7], "ylim": [0.0, 5e-8], "zlim": [1e-8, 2.8e-7]}
    for m, p, o, dt in ((mc, pos, ori, 1e-13), (golden.mc, golden.pos_nm, golden.orient, golden.dt), (mc, pos, ori, 7e-13)):
        res = []
        for on in (False, True):
            e = Emul(m)
            e.kubo_init(p, o)
            e.set_runs(on)
            e.create_philox(300, seed=5)
            msd = e.kubo_step(dt, 120, trace_cap=1 << 12)
            res.append((e.particles(), msd, e.trace(), e.hops()))
        (pa, ma, (oa, fa), ha), (pb, mb, (ob, fb), hb) = res
        assert all(np.array_equal(pa[k], pb[k]) for k in pa) and np.array_equal(ma, mb)
        assert np.array_equal(oa, ob) and np.array_equal(fa, fb) and ha == hb and ha > 300


def test_division_by_constants_is_the_ieee_division():
    """div_by (reciprocal, exact remainder, correction) against x / c: the draws themselves, rate x draw products and
    distances, for RAND_MAX and a set of velocities."""
    import ctypes
    e = Emul(base_mc())
    f = e.L.emul_div_by_mismatches
    f.restype = ctypes.c_int64
    f.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_double]
    rng = np.random.default_rng(99)
    draws = np.concatenate([np.arange(0, 1 << 16), (1 << 31) - 1 - np.arange(0, 1 << 16), rng.integers(0, 1 << 31, size=3_000_000)]).astype(np.float64)
    assert f(draws.ctypes.data, len(draws), 2147483647.0) == 0
    prod = np.ldexp(1.0 + rng.random(5_000_000), rng.integers(10, 90, size=5_000_000).astype(np.int32))
    assert f(prod.ctypes.data, len(prod), 2147483647.0) == 0
    dist = np.ldexp(1.0 + rng.random(2_000_000), -rng.integers(20, 50, size=2_000_000).astype(np.int32))
    for v in (2e5, 1e5, 3.3e5, 123456.789, 1e4, 7.7e5):
        assert f(dist.ctypes.data, len(dist), v) == 0


def test_shortcuts_on_random_trim_boxes():
    """Random small films cut by random trim boxes (the swap-with-last trim permutes sites, so runs of memory-consecutive
    chain neighbours break in arbitrary places): segment-time walks + stored directions + top entries change nothing."""
    rng = np.random.default_rng(2024)
    tried = 0
    for case in range(12):
        pos, ori = film.film(NT=int(rng.integers(20, 60)), NP=int(rng.integers(8, 40)), a=float(rng.choice([2.0, 5.0])),
                             LX=float(rng.uniform(60, 160)), LY=float(rng.uniform(20, 60)), seed=int(rng.integers(1, 1 << 30)))
        lo, hi = pos.reshape(3, -1).min(axis=1) * 1e-9, pos.reshape(3, -1).max(axis=1) * 1e-9
        cut = rng.uniform(0.0, 0.3, size=(3, 2)) * (hi - lo)[:, None]
        mc = dict(base_mc())
        mc["trim limits"] = {k: [float(lo[i] + cut[i, 0]), float(hi[i] - cut[i, 1])] for i, k in enumerate(("xlim", "ylim", "zlim"))}
        states = []
        try:
            for on in (False, True):
                e = Emul(mc)
                e.kubo_init(pos, ori)
                if len(e.inject()) == 0 or not np.all(np.diff(e.csr()[0]) > 0):
                    raise ValueError("empty injection region or a site without neighbours")  # the engine refuses these
                e.set_runs(on)
                e.set_top_entries(on)
                e.create_philox(150, seed=case + 1)
                msd = e.kubo_step(1e-13 if case % 2 else 4e-13, 60, trace_cap=1 << 12)
                states.append((e.particles(), msd, e.trace(), e.hops()))
        except (ValueError, AssertionError):
            continue  # e.g. an empty injection region or an isolated site after the cut: not what this test is about
        tried += 1
        (pa, ma, (oa, fa), ha), (pb, mb, (ob, fb), hb) = states
        assert all(np.array_equal(pa[k], pb[k]) for k in pa) and np.array_equal(ma, mb), case
        assert np.array_equal(oa, ob) and np.array_equal(fa, fb) and ha == hb, case
    assert tried >= 6


def test_guided_search_equals_reference_loop_for_every_bucket():
    e = Emul(base_mc())
    rng = np.random.default_rng(5)
    e.L.emul_select_guided.restype = __import__("ctypes").c_int64
    e.L.emul_select_guided.argtypes = [__import__("ctypes").c_void_p, __import__("ctypes").c_int64, __import__("ctypes").c_int32]
    for _ in range(1500):
        d = int(rng.integers(1, 300)) if _ % 3 else int(rng.integers(250, 3000))  # long rows store scaled guide entries
        c = np.ascontiguousarray(np.cumsum(rng.random(d) ** 8 * (rng.random(d) > 0.2) * 10.0 ** rng.integers(8, 15)))
        if c[-1] == 0:
            continue
        draws = np.concatenate([rng.integers(0, 2 ** 31, 24), np.arange(16) << 27, (np.arange(1, 17) << 27) - 1, [0, 2 ** 31 - 1]])
        for r in draws:
            r = int(min(r, 2 ** 31 - 1))
            dice = c[-1] * float(r) / 2147483647.0
            assert e.L.emul_select_guided(c.ctypes.data, d, r) == T1m.select(c, dice)


def test_engine_core_uses_the_same_philox_stream_as_the_oracle():
    import ctypes
    e = Emul(base_mc())
    out = (ctypes.c_uint32 * 2)()
    e.L.emul_philox2x32.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    e.L.emul_philox2x32(0x243F6A88, 0x85A308D3, 0x13198A2E, out)
    assert [hex(x) for x in out] == ["0xdd7ce038", "0xf62a4c12"]


def test_untrimmed_bigger_film_against_oracle():
    pos, ori = film.film(NT=120, NP=60, a=5.0, LX=300.0, LY=80.0, seed=5)
    mc = base_mc()
    e = Emul(mc)
    e.kubo_init(pos, ori)
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    for a, b in zip(e.csr(), t.csr()):
        assert np.array_equal(a, b)
    t.set_memo(True)
    t.draws_philox(3)
    t.create_particles(500)
    t.kubo_step(1e-13, 50, want_msd=False)
    e.create_philox(500, seed=3)
    e.kubo_step(1e-13, 50)
    pe, pt = e.particles(), t.particles()
    for k in STATE_KEYS:
        assert np.array_equal(pe[k], pt[k]), k


def test_table_logarithm_is_within_one_ulp_of_glibc_on_every_possible_draw(tmp_path):
    """fast_log_unit (csrc/fast_log.h) replaces the library logarithm in the free-flight draw: compared with glibc's log on
    ALL 2^31 - 1 arguments r / RAND_MAX (same IEEE arithmetic on host and device, so the CPU run covers the kernel), and
    div_by(r, RAND_MAX) with the division on all of them."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "log_exh")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-fopenmp", "-x", "c++",
                           os.path.join(root, "tools", "log_exhaustive.c"), "-o", exe])
    out = json.loads(subprocess.run([exe], capture_output=True, text=True, check=True).stdout)
    assert out["args"] == 2147483647 and out["more"] == 0 and out["div_by_mismatch"] == 0
    assert out["one_ulp"] < 0.002 * out["args"]      # 99.9 % of the results are glibc's bits, the rest one ulp away


def test_windowed_grid_search_equals_the_reference_scan():
    """scattering_struct::get_rate picks arma::abs(grid - x).index_min() (scattering_struct.h:42-49).  The table build
    looks only near round((x - start)/step) on evenly spaced grids: same index for ties (exact midpoints: the first one
    wins), grid values, values beyond the ends, NaN and infinities; irregular grids fall back to the full scan."""
    import emul
    rng = np.random.default_rng(0)
    for lo, hi, n in ((0.0, 180 * 3.141592 / 180, 21), (1.5e-9, 10e-9, 11), (-10e-9, 10e-9, 11), (-8e-9, 8e-9, 9), (0.0, 1.0, 2), (5.0, 5.0, 1)):
        grid = lo + (hi - lo) * np.arange(n) / max(1, n - 1) if n > 1 else np.array([lo])
        step = (hi - lo) / max(1, n - 1)
        mids = 0.5 * (grid[:-1] + grid[1:])
        xs = np.concatenate([grid, mids, np.nextafter(mids, -np.inf), np.nextafter(mids, np.inf), np.nextafter(grid, np.inf),
                             rng.uniform(lo - 3 * step - 1e-9, hi + 3 * step + 1e-9, 200000),
                             [np.nan, np.inf, -np.inf, 0.0, -0.0, 1e300, -1e300, lo - 1e-30, hi + 1e-30, 1.0, -1.0, 7.3e-2, 1e7 * step, -1e7 * step],
                             hi + step * 10.0 ** rng.uniform(0, 17, 2000), lo - step * 10.0 ** rng.uniform(0, 17, 2000)])
        bad, hinted = emul.argmin_mismatches(grid, xs)
        assert bad == 0 and hinted == (n >= 2 and hi > lo)
    irregular = np.array([0.0, 0.1, 0.5, 0.55, 2.0])
    bad, hinted = emul.argmin_mismatches(irregular, rng.uniform(-1, 3, 10000))
    assert bad == 0 and not hinted
