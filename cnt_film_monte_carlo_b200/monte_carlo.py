"""Host-side mirror of the reference's ``mc::monte_carlo`` class (src/monte_carlo/monte_carlo.h:34-842) over the C ABI.

Same method names, argument meaning, error behaviour, directory handling and output-file formats as the reference, so
``main()`` of the reference (src/main.cpp:18-114) reads the same when written against this class -- see
:func:`main`.  All computation happens in libcntmc.so on the GPU; this file only does what the reference's host code
does around it (directories, JSON, text files).  The C++ twin of this file is cpp/monte_carlo.hpp.
"""
from __future__ import annotations

import json
import os
import shutil
import sys
from typing import Optional

import numpy as np

from .engine import CntmcError, Engine


def _expand(path: str) -> str:
    """prepare_directory.hpp:12-16: only a leading '~' is expanded."""
    if path.startswith("~"):
        return os.environ.get("HOME", "") + path[1:]
    return path


def prepare_directory(path: str, keep_old_files: bool = True) -> str:
    """helper/prepare_directory.hpp:8-65: create the output directory; rotate (dir -> dir.N) or delete a non-empty one."""
    path = _expand(path)
    if not os.path.exists(path):
        os.makedirs(path)
        return path
    if not os.path.isdir(path):
        raise ValueError("The input value for output directory is not acceptable.")
    if os.listdir(path):
        if keep_old_files:
            count = 1
            while os.path.exists(f"{path}.{count}"):
                count += 1
            os.rename(path, f"{path}.{count}")
        else:
            shutil.rmtree(path)
        os.makedirs(path)
    return path


def check_directory(path: str, should_be_empty: bool = False) -> str:
    """helper/prepare_directory.hpp:68-105."""
    path = _expand(path)
    if not os.path.exists(path):
        raise ValueError("directory does NOT exists!!!")
    if not os.path.isdir(path):
        raise ValueError("input path is NOT a directory!!!")
    if not os.listdir(path) and not should_be_empty:
        raise ValueError("directory is empty!!!")
    return path


def _sci(v: float) -> str:
    """std::showpos << std::scientific (6 digits), the format of every number the reference writes."""
    return "%+.6e" % v


class monte_carlo:
    """``mc::monte_carlo``.  ``j`` is the "exciton monte carlo" block of input.json (monte_carlo.h:116-136)."""

    def __init__(self, j: dict, device: Optional[int] = None, seed: int = 100, quiet: bool = False):
        self._json_prop = dict(j)
        self._quiet = quiet
        keep = bool(j.get("keep old results", True))
        self._output_directory = prepare_directory(j["output directory"], keep)
        self._input_directory = check_directory(j["mesh input directory"], False)
        self._engine = Engine(self._json_prop, device=device)
        self._seed = seed  # main.cpp:30 seeds glibc with 100; here it keys the counter-based exciton streams
        self._msd_file = None
        self._pop_file = None
        self._curr_file = None
        self._last_msd = None
        self._last_metrics = None
        self._n_steps_done = 0

    # -- accessors (monte_carlo.h:139-154) --------------------------------------------------------------------------
    def time(self) -> float:
        return self._engine.time()

    def output_path(self) -> str:
        return self._output_directory

    def input_path(self) -> str:
        return self._input_directory

    def number_of_particles(self) -> int:
        return self._engine.number_of_particles()

    def kubo_max_time(self) -> float:
        return self._engine.kubo_max_time()

    @property
    def engine(self) -> Engine:
        return self._engine

    def save_json_properties(self) -> None:
        """monte_carlo.h:319-324"""
        with open(os.path.join(self._output_directory, "input.json"), "w") as f:
            json.dump(self._json_prop, f, indent=4)
            f.write("\n")

    def _say(self, *a):
        if not self._quiet:
            print(*a)

    # -- Green-Kubo flavour ---------------------------------------------------------------------------------------------
    def kubo_init(self) -> None:
        """monte_carlo.cpp:254-305"""
        self._say("maximum hopping radius: %g [nm]" % (float(self._json_prop["max hopping radius [m]"]) * 1e9))
        self._say("exciton velocity [m/s]: %g" % float(self._json_prop["exciton velocity [m/s]"]))
        self._engine.load_mesh(self._input_directory)
        self._engine.kubo_init()
        if self._json_prop.get("rate type") == "davoody":
            self.save_scat_table()  # create_davoody_scatt_table leaves its table in the output directory (monte_carlo.cpp:150)
        d = self._engine.domain() * 1e9
        self._say("\nsimulation domain AFTER trimming:\n    x (%+f , %+f) [nm]\n    y (%+f , %+f) [nm]\n    z (%+f , %+f) [nm]\n"
                  % (d[0], d[3], d[1], d[4], d[2], d[5]))
        self._say("total number of scatterers: %d" % self._engine.num_sites())

    def kubo_create_particles(self) -> None:
        """monte_carlo.cpp:308-316"""
        self._engine.kubo_create_particles(0, seed=self._seed)

    def kubo_step(self, dt: float) -> None:
        """monte_carlo.cpp:319-342"""
        self._last_msd = self._engine.kubo_step(dt, 1)[0]
        self._n_steps_done += 1

    def kubo_save_avg_dispalcement_squared(self) -> None:  # [sic] the reference's spelling
        """monte_carlo.cpp:382-409: one row  time,<dx^2>,<dy^2>,<dz^2>  per call."""
        self._write_msd_rows(np.array([self.time()]), self._last_msd[None, :])

    def kubo_run(self, dt: float, nsteps: int) -> np.ndarray:
        """``nsteps`` x { kubo_step(dt); kubo_save_avg_dispalcement_squared() } in one engine call (same rows)."""
        t0 = self.time()
        msd = self._engine.kubo_step(dt, nsteps)
        times = np.empty(nsteps)
        t = t0
        for s in range(nsteps):  # _time += dt, one addition per step (monte_carlo.cpp:341)
            t += dt
            times[s] = t
        self._write_msd_rows(times, msd)
        self._last_msd = msd[-1]
        self._n_steps_done += nsteps
        return msd

    def _write_msd_rows(self, times, msd) -> None:
        if self._msd_file is None:
            self._msd_file = open(os.path.join(self._output_directory, "particle_dispalcement.avg.squared.dat"), "w")
            self._msd_file.write("# this file contains the average of dx^2, dy^2, and dz^2 of the particle ensemble over time\n"
                                 "# number of particles: %d\n\ntime,x,y,z\n" % self.number_of_particles())
        for t, row in zip(times, msd):
            self._msd_file.write(",".join(_sci(v) for v in (t, row[0], row[1], row[2])) + "\n")
        self._msd_file.flush()

    def kubo_save_individual_particle_dispalcements(self) -> None:
        """monte_carlo.cpp:345-380: three files, one column per exciton."""
        delta = self._engine.particles()["delta"]
        for c, ax in enumerate("xyz"):
            path = os.path.join(self._output_directory, f"particle_dispalcement.{ax}.dat")
            new = not os.path.exists(path) or self._n_steps_done <= 1
            with open(path, "w" if new else "a") as f:
                if new:
                    f.write("time" + "".join(",%+d" % i for i in range(delta.shape[1])) + "\n")
                f.write(_sci(self.time()) + "".join("," + _sci(v) for v in delta[c]) + "\n")

    # -- contact flavour ----------------------------------------------------------------------------------------------------
    def init(self, c1_pop: int = 1100, c2_pop: int = 0) -> None:
        """monte_carlo.h:157-195 (the reference hard-codes the contact populations 1100 and 0, :191-192)."""
        self._engine.load_mesh(self._input_directory)
        self._engine.init(c1_pop, c2_pop, seed=self._seed)
        if self._json_prop.get("rate type") == "davoody":
            self.save_scat_table()  # monte_carlo.cpp:150
        self._n_seg = self._engine.number_of_segments()
        self._area = self._engine.area()
        self._domain = self._engine.domain()
        self._say("number of segments: %d" % self._n_seg)
        self._say("total number of scatterers: %d" % self._engine.num_sites())
        self.get_scatterer_statistics()

    def get_scatterer_statistics(self) -> None:
        """monte_carlo.h:691-719: scatterer_statistics.dat (called from init(), :183)."""
        ymin, ymax = self._domain[1], self._domain[4]
        dy = (ymax - ymin) / self._n_seg
        pop = self._engine.scatterer_statistics()
        total = self._engine.num_sites()
        with open(os.path.join(self._output_directory, "scatterer_statistics.dat"), "w") as f:
            f.write("position,distribution,population,density\n")
            for i, (n, a) in enumerate(zip(pop, self._area)):
                f.write("%.6e,%.6e,%d,%.6e\n" % (ymin + (i + 0.5) * dy, float(n) / float(total), n, float(n) / (a * dy)))

    def track_particle(self, dt: float, file_no: int, max_steps: int = 1 << 20) -> bool:
        """monte_carlo.h:786-818: particle_path.<file_no>.dat; the exciton's stream is (seed, 2^63 | file_no), a
        domain of its own (never the draws of exciton file_no of the population).  Returns whether
        the last slab was reached within max_steps (the reference's loop is unbounded)."""
        path, reached = self._engine.track_particle(dt, seed=self._seed, global_id=(1 << 63) | file_no, max_steps=max_steps)
        with open(os.path.join(self._output_directory, "particle_path.%d.dat" % file_no), "w") as f:
            for r in path:
                f.write("   %+.6e %+.6e %+.6e\n" % tuple(r))
            f.write("\n")
        return reached

    def save_scat_table(self) -> None:
        """scattering_struct::save (scattering_struct.h:56-94) into the output directory (monte_carlo.cpp:150)."""
        self._engine.save_rate_table(self._output_directory)

    def step(self, dt: float) -> None:
        """monte_carlo.h:343-355 -- together with save_metrics/repopulate_contacts (the engine fuses the three)."""
        pop, cur = self._engine.step(dt, 1)
        self._last_metrics = (pop[0], cur[0], dt)

    def save_metrics(self, dt: float) -> None:
        """monte_carlo.h:519-522"""
        pop, cur, _ = self._last_metrics
        self._write_population(pop)
        self._write_currents(cur, dt)

    def repopulate_contacts(self) -> None:
        """monte_carlo.h:443-455: done on the device at the end of step(); kept for call-site compatibility."""

    def run(self, dt: float, nsteps: int):
        """``nsteps`` x { step; save_metrics; repopulate_contacts } in one engine call."""
        pop, cur = self._engine.step(dt, nsteps)
        return pop, cur

    def _write_population(self, pop) -> None:
        """monte_carlo.h:525-581"""
        ymin, ymax = self._domain[1], self._domain[4]
        dy = (ymax - ymin) / self._n_seg
        if self._pop_file is None:
            f = self._pop_file = open(os.path.join(self._output_directory, "population_profile.dat"), "w")
            f.write("area" + "".join("," + _sci(a) for a in self._area) + "\n\n")
            f.write("dy" + "".join("," + _sci(dy) for _ in self._area) + "\n\n")
            f.write("section pos" + "".join("," + _sci(ymin + (i + 0.5) * dy) for i in range(self._n_seg)) + "\n\n")
            f.write("time" + "".join(",section%d" % i for i in range(self._n_seg)) + "\n")
        self._pop_file.write(_sci(self.time()) + "".join("," + _sci(p / (a * dy)) for p, a in zip(pop, self._area)) + "\n")
        self._pop_file.flush()

    def _write_currents(self, cur, dt) -> None:
        """monte_carlo.h:584-643"""
        ymin, ymax = self._domain[1], self._domain[4]
        n = self._n_seg
        dy = (ymax - ymin) / n
        area_if = [(self._area[i - 1] + self._area[i]) / 2 for i in range(1, n)]
        if self._curr_file is None:
            f = self._curr_file = open(os.path.join(self._output_directory, "region_current.dat"), "w")
            f.write("interface area" + "".join("," + _sci(a) for a in area_if) + "\n\n")
            f.write("interface pos" + "".join("," + _sci(ymin + dy * i) for i in range(1, n)) + "\n\n")
            f.write("time" + "".join(",interface%+d" % (i - 1) for i in range(1, n)) + "\n")  # showpos is still on in the reference
        self._curr_file.write(_sci(self.time()) + "".join("," + _sci(c / (a * dt)) for c, a in zip(cur, area_if)) + "\n")
        self._curr_file.flush()

    def close(self) -> None:
        for f in (self._msd_file, self._pop_file, self._curr_file):
            if f is not None:
                f.close()
        self._engine.close()


def main(argv=None) -> int:
    """src/main.cpp:18-80: the Green-Kubo driver.  ``python -m cnt_film_monte_carlo_b200.monte_carlo input.json``"""
    argv = sys.argv if argv is None else argv
    filename = argv[1] if len(argv) > 1 else "input.json"
    with open(filename) as f:
        j = json.load(f)
    if "exciton monte carlo" not in j:
        raise ValueError('json input file does not contain "exciton monte carlo"')
    json_mc = j["exciton monte carlo"]
    if json_mc.get("rate type") == "davoody":
        json_mc["cnts"] = j.get("cnts")
    time_step = float(json_mc["monte carlo time step"])
    sim = monte_carlo(json_mc)
    sim.kubo_init()
    sim.save_json_properties()
    sim.kubo_create_particles()
    batch = int(json_mc.get("steps per engine call", 1024))  # optional key; the rows written are the same
    while sim.time() < sim.kubo_max_time():
        # number of steps the reference's loop would still take, computed with its own floating-point accumulation
        t, n = sim.time(), 0
        while t < sim.kubo_max_time() and n < batch:
            t += time_step
            n += 1
        sim.kubo_run(time_step, n)
        print("kubo simulation: current time [seconds]: %e .... max time [seconds]: %e" % (sim.time(), sim.kubo_max_time()), end="\r")
    print("\nGreen-Kubo simulation finished!")
    sim.close()
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except (CntmcError, ValueError, KeyError) as e:
        sys.exit("error: %s" % e)
