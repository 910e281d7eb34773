"""The oracle on BASELINE config 1: input.json's own trim limits and dt = 1e-15 s, pinned to what the
reference's code computed (tests/golden/c1_input_json.npz)."""
import numpy as np

import c1_case
from oracle import t1 as T1m


def test_oracle_reproduces_the_reference_on_config_1():
    mc, z, pos, ori = c1_case.load()
    t = T1m.T1()
    t.kubo_init(mc, pos, ori)
    row_ptr, nbr, cum = t.csr()
    c1_case.check_setup(z, t.sites(), t.domain(), t.removal_domain(), t.inject(), row_ptr, nbr, cum)
    P = len(z["draw_off"]) - 1
    t.draws_replay(z["draw_off"], z["draws"])
    t.create_particles(P)
    p0 = t.particles()
    for k in ("site", "pos", "ff", "heading"):
        assert np.array_equal(p0[k], z["p0_" + k]), k
    msd = t.kubo_step(float(z["dt"]), int(z["nsteps"]))           # 20 000 steps of 1e-15 s
    p1 = t.particles()
    for k in ("site", "pos", "delta", "ff", "heading"):
        assert np.array_equal(p1[k], z["p1_" + k]), k
    assert np.array_equal(msd[c1_case.MSD_STRIDE - 1::c1_case.MSD_STRIDE], z["msd_sample"])
    assert not t.replay_exhausted()
