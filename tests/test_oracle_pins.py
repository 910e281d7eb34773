"""Known answers that pin the oracle independently of any reference run (SURVEY.md §8c pins 1-5, 8)."""
import ctypes
import math

import numpy as np
import pytest

from oracle import t1 as T1m


def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10
    assert [hex(x) for x in T1m.philox4x32_10([0] * 4, [0] * 2)] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in T1m.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in T1m.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == [
        "0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_philox2x32_known_answers():
    # Random123 kat_vectors for philox2x32-10: the generator behind the engine's exciton streams
    assert [hex(x) for x in T1m.philox2x32_10([0, 0], 0)] == ["0xff1dae59", "0x6cd10df2"]
    assert [hex(x) for x in T1m.philox2x32_10([0xFFFFFFFF, 0xFFFFFFFF], 0xFFFFFFFF)] == ["0x2c3f628b", "0xab4fd7ad"]
    assert [hex(x) for x in T1m.philox2x32_10([0x243F6A88, 0x85A308D3], 0x13198A2E)] == ["0xdd7ce038", "0xf62a4c12"]


def test_philox_draw_layout():
    # draw k of exciton g = word k&1 of block k>>1 of the stream keyed by (seed, g), shifted to 31 bits
    for seed, g in ((7, 12345), (0x1234567890ABCDEF, 0x0000000500000007)):
        sh, gh = seed >> 32, g >> 32
        key = (seed & 0xFFFFFFFF) ^ (((sh << 16) | (sh >> 16)) & 0xFFFFFFFF) ^ ((gh * 0x9E3779B9) & 0xFFFFFFFF)
        for k in (0, 1, 2, 3, 4, 9, 1023):
            blk = T1m.philox2x32_10([k >> 1, g & 0xFFFFFFFF], key)
            assert T1m.philox_draw(seed, g, k) == int(blk[k & 1]) >> 1
    assert T1m.philox_draw(7, 12345, 0) == int(T1m.philox2x32_10([0, 12345], 7)[0]) >> 1  # small seeds: key == seed


def test_glibc_srand100_first_draws():
    # main.cpp:30 seeds with 100; glibc's TYPE_3 additive generator (SURVEY.md §8c pin 8)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(100)
    assert [libc.rand() for _ in range(3)] == [677741240, 611911301, 516687479]


def test_linspace_matches_armadillo_rule():
    x = T1m.linspace(1.5e-9, 10e-9, 11)
    d = (10e-9 - 1.5e-9) / 10
    assert x[0] == 1.5e-9 and x[-1] == 10e-9
    assert all(x[i] == 1.5e-9 + i * d for i in range(10))


def test_forster_entry_by_hand():
    mc = {"rate type": "forster", "zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11],
          "axis shift 2 [m]": [-10e-9, 10e-9, 11], "theta [degrees]": [0, 180, 21]}
    t = T1m.forster_table(mc)
    assert t["rates"].shape == (21, 11, 11, 11)
    # theta grid uses the reference's truncated pi (helper/constants.h:10)
    assert t["theta"][-1] == 180 * (3.141592 / 180)
    # monte_carlo.cpp:183-187 for theta = 90 deg (index 10), z = 1.5 nm, a1 = +2 nm (6), a2 = -4 nm (3)
    th, z, a1, a2 = t["theta"][10], t["z"][0], t["a1"][6], t["a2"][3]
    r1 = np.array([a1, 0, 0])
    r2 = np.array([a2 * math.cos(th), a2 * math.sin(th), z])
    dR = r1 - r2
    u = lambda v: v / np.linalg.norm(v)
    af = math.cos(th) - 3 * np.dot(u(r1), u(dR)) * np.dot(u(r2), u(dR))
    expect = 1e15 * af ** 2 * (1e-9 / np.linalg.norm(dR)) ** 6
    assert t["rates"][10, 0, 6, 3] == pytest.approx(expect, rel=1e-12)
    # a1 = 0 makes normalise(r1) the zero vector: angle factor is cos(theta) alone
    th0, a20 = t["theta"][4], t["a2"][8]
    dR0 = np.linalg.norm([0 - a20 * math.cos(th0), -a20 * math.sin(th0), -t["z"][2]])
    assert t["rates"][4, 2, 5, 8] == pytest.approx(1e15 * math.cos(th0) ** 2 * (1e-9 / dR0) ** 6, rel=1e-12)
    wong = T1m.forster_table(dict(mc, **{"rate type": "wong"}))
    assert np.allclose(wong["rates"] * 100, t["rates"], rtol=1e-15)


def test_get_rate_is_nearest_grid_point_first_minimum_wins():
    s = T1m.T1()
    grid = np.array([0.0, 1.0, 2.0])
    rates = np.arange(81, dtype=float).reshape(3, 3, 3, 3)
    s.set_table(dict(theta=grid, z=grid, a1=grid, a2=grid, rates=rates))
    assert s.get_rate(0.4, 1.6, -5.0, 9.0) == rates[0, 2, 0, 2]  # clamps outside the grid
    assert s.get_rate(0.5, 0.0, 0.0, 0.0) == rates[0, 0, 0, 0]   # exact tie -> first (lower) index
    assert s.get_rate(float("nan"), 1.0, 1.0, 1.0) == rates[0, 1, 1, 1]  # NaN never wins -> index 0


def test_select_literal_semantics():
    cum = np.array([1.0, 3.0, 6.0])
    pick = lambda dice: T1m.select(cum, dice)
    assert [pick(0.0), pick(0.999), pick(1.0), pick(2.999), pick(3.0), pick(5.999)] == [0, 0, 1, 1, 2, 2]
    assert pick(6.0) == 2  # dice == total selects the last neighbour
    assert T1m.select(np.array([4.2]), 1.0) == 0
    # zero-rate entries (flat stretches) can never be chosen unless they are last
    cum0 = np.array([0.0, 0.0, 2.0, 2.0, 5.0, 5.0])
    assert [T1m.select(cum0, d) for d in (0.0, 1.0, 2.0, 4.9, 5.0)] == [2, 2, 4, 4, 5]
    # brute force: first k with cum[k] > dice, else the last
    rng = np.random.default_rng(3)
    for _ in range(2000):
        d = int(rng.integers(1, 40))
        c = np.cumsum(rng.random(d) * (rng.random(d) > 0.2))
        dice = c[-1] * rng.random() if rng.random() > 0.1 else rng.choice(c)
        above = np.nonzero(c > dice)[0]
        assert T1m.select(c, dice) == (above[0] if len(above) else d - 1)
