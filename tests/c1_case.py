"""BASELINE config 1 (the reference's own input.json on the stand-in film) as a shared test case: the fixture
tests/golden/c1_input_json.npz was computed by the reference's own code (tests/golden/make_golden_c1.py)."""
import json
import os

import numpy as np

from cnt_film_monte_carlo_b200 import film

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROW_STRIDE, MSD_STRIDE = 41, 50


def load():
    with open(os.path.join(GOLDEN_DIR, "c1_input_json.json")) as f:
        mc = json.load(f)["exciton monte carlo"]
    z = np.load(os.path.join(GOLDEN_DIR, "c1_input_json.npz"))
    pos, ori = film.film(**film.CONFIG_FILMS["C1"])
    return mc, z, pos, ori


def checksums(row_ptr, nbr, cum):
    k = np.arange(len(nbr), dtype=np.uint64)
    return np.array([np.sum(nbr.astype(np.uint64) * (k % np.uint64(1000003) + np.uint64(1)), dtype=np.uint64),
                     np.sum(np.ascontiguousarray(cum).view(np.uint64) ^ (k * np.uint64(0x9E3779B97F4A7C15)), dtype=np.uint64)], np.uint64)


def check_setup(z, sites, domain, removal, inject, row_ptr, nbr, cum):
    """`sites` etc. as produced by the oracle or the engine; everything must be the reference's bits."""
    N = int(z["n_sites"])
    assert len(sites["left"]) == N and N < 20000                      # input.json's ylim = [0, 100 nm] cuts tilted tubes
    assert np.array_equal(sites["left"], z["site_left"]) and np.array_equal(sites["right"], z["site_right"])  # the swap-loop permutation
    assert np.array_equal(sites["pos"][:, ::ROW_STRIDE], z["site_pos_sample"])
    assert np.array_equal(sites["max_rate"][::ROW_STRIDE], z["site_max_rate_sample"])
    assert np.array_equal(domain, z["domain"]) and np.array_equal(removal, z["removal"]) and np.array_equal(inject, z["inject"])
    assert np.array_equal(np.diff(row_ptr).astype(np.uint16), z["degree"])
    assert np.array_equal(checksums(row_ptr, nbr, cum), z["csr_checksums"])   # every entry of every row
    rows = np.arange(0, N, ROW_STRIDE)
    assert np.array_equal(np.concatenate([nbr[row_ptr[i]:row_ptr[i + 1]] for i in rows]), z["row_sample_nbr"])
    assert np.array_equal(np.concatenate([cum[row_ptr[i]:row_ptr[i + 1]] for i in rows]), z["row_sample_cum"])
