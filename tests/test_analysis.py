"""The output-file readers / fits against files written by the reference's own code (golden fixture)."""
import numpy as np

from cnt_film_monte_carlo_b200 import analysis


def write(tmp_path, g):
    for key, name in (("contact_pop_file", "population_profile.dat"), ("contact_curr_file", "region_current.dat"),
                      ("contact_stat_file", "scatterer_statistics.dat"), ("ref_program_output", "particle_dispalcement.avg.squared.dat")):
        with open(tmp_path / name, "wb") as f:
            f.write(bytes(g.z[key]))
    return str(tmp_path)


def test_reads_the_reference_programs_msd_file(golden_small, tmp_path):
    g = golden_small
    d = write(tmp_path, g)
    m = analysis.read_msd(d)
    n = len(m["time"])
    assert n == g.z["nsteps"]
    assert np.allclose(np.stack([m["x"], m["y"], m["z"]], 1), g.z["msd"][:n], rtol=2e-6)   # 7 printed digits
    assert np.allclose(m["time"], g.z["dt"] * np.arange(1, n + 1), rtol=1e-6)
    k = analysis.kubo_diffusion(d, skip_fraction=0.2)
    for ax in "xyz":
        assert k["D"][ax] > 0 and np.isfinite(k["msd_over_time"][ax]).all()
    # a pure diffusion law is fitted exactly
    t = m["time"]
    with open(tmp_path / "particle_dispalcement.avg.squared.dat", "w") as f:
        f.write("# c\n# number of particles: 1\n\ntime,x,y,z\n" + "".join("%+e,%+e,%+e,%+e\n" % (x, 2 * 3e-4 * x, 2 * 1e-4 * x, 0) for x in t))
    k = analysis.kubo_diffusion(d)
    assert np.isclose(k["D"]["x"], 3e-4, rtol=1e-5) and np.isclose(k["D"]["y"], 1e-4, rtol=1e-5) and abs(k["D"]["z"]) < 1e-12


def test_reads_the_reference_contact_files(golden_small, tmp_path):
    g = golden_small
    d = write(tmp_path, g)
    n_seg = int(g.mc["number of segments"])
    p, c, s = analysis.read_population(d), analysis.read_current(d), analysis.read_scatterer_stats(d)
    steps = len(g.z["contact_num_particles"]) - 1
    assert p["pop"].shape == (steps, n_seg) and c["current"].shape == (steps, n_seg - 1)
    assert np.allclose(p["area"], g.z["contact_area"], rtol=2e-6)
    assert np.allclose(c["area"], (g.z["contact_area"][:-1] + g.z["contact_area"][1:]) / 2, rtol=2e-6)
    assert len(p["dy"]) == n_seg and np.allclose(np.diff(p["pos"]), p["dy"][0], rtol=1e-5)
    assert np.allclose(c["pos"], p["pos"][:-1] + p["dy"][0] / 2, rtol=1e-5)
    assert list(s) == ["position", "distribution", "population", "density"] and s["population"].sum() == len(g.z["site_left"])
    assert np.isclose(s["distribution"].sum(), 1.0, rtol=1e-5)
    f = analysis.contact_diffusion(d)
    assert f["diffusion"].shape == (n_seg - 1,) and np.allclose(f["steady_current"], c["current"].mean(axis=0))
    assert analysis.main([d, "--kubo", "--diffusion"]) == 0
