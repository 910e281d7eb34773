"""Readers and fits for the simulation's output files (the numerical part of the reference's
``visualization/monte_carlo_results.py``, without plotting).

Files (written by the reference and, in the same format, by this package's ``monte_carlo`` mirror / ``cntmc_main``):

* ``particle_dispalcement.avg.squared.dat``  -- Green-Kubo run: ``time, <dx^2>, <dy^2>, <dz^2>`` (monte_carlo.cpp:382-409)
* ``population_profile.dat``                 -- contact run: exciton density per slab (monte_carlo.h:525-581)
* ``region_current.dat``                     -- contact run: current density per interface (monte_carlo.h:584-643)
* ``scatterer_statistics.dat``               -- sites per slab (monte_carlo.h:691-719)

    python -m cnt_film_monte_carlo_b200.analysis <output directory> --kubo [--skip 0.2]
    python -m cnt_film_monte_carlo_b200.analysis <output directory> --diffusion
"""
from __future__ import annotations

import argparse
import json
import os
from typing import Dict

import numpy as np


def _rows(path: str, skip: int):
    with open(path) as f:
        return f.read().splitlines()[skip:]


def _numbers(line: str, drop_label: bool = True):
    cells = line.split(",")
    return np.array([float(c) for c in (cells[1:] if drop_label else cells)])


def read_msd(directory: str) -> Dict[str, np.ndarray]:
    """``particle_dispalcement.avg.squared.dat``: two comment lines, a blank line, the header ``time,x,y,z``, then rows
    (monte_carlo_results.py:478-483 reads it with ``skiprows=3``)."""
    lines = _rows(os.path.join(directory, "particle_dispalcement.avg.squared.dat"), 3)
    assert lines[0].strip() == "time,x,y,z", lines[0]
    data = np.array([_numbers(ln, drop_label=False) for ln in lines[1:] if ln.strip()])
    return dict(time=data[:, 0], x=data[:, 1], y=data[:, 2], z=data[:, 3])


def kubo_diffusion(directory: str, skip_fraction: float = 0.0) -> Dict[str, object]:
    """Diffusion coefficients from the mean squared displacement: ``<d_a^2>(t) = 2 D_a t``.

    ``D`` is the least-squares slope / 2 over the rows after the first ``skip_fraction`` of the run (the ballistic
    transient); ``msd_over_time`` is the curve the reference plots (monte_carlo_results.py:489-496: ``<d^2>/t``, which
    tends to ``2 D``)."""
    m = read_msd(directory)
    t = m["time"]
    first = int(len(t) * skip_fraction)
    out: Dict[str, object] = {"time": t, "msd_over_time": {}, "D": {}}
    for ax in "xyz":
        out["msd_over_time"][ax] = m[ax] / t
        tt, yy = t[first:], m[ax][first:]
        if len(tt) >= 2:
            slope = np.polyfit(tt, yy, 1)[0]
        else:
            slope = yy[-1] / tt[-1]
        out["D"][ax] = 0.5 * slope
    total = (m["x"] + m["y"] + m["z"]) / 3
    out["msd_over_time"]["total"] = total / t
    out["D"]["total"] = 0.5 * (np.polyfit(t[first:], total[first:], 1)[0] if len(t) - first >= 2 else total[-1] / t[-1])
    return out


def read_current(directory: str) -> Dict[str, np.ndarray]:
    """``region_current.dat`` (monte_carlo_results.py:31-68): line 0 interface areas, line 2 interface positions, line 4
    the column names, then ``time, current density per interface``."""
    lines = _rows(os.path.join(directory, "region_current.dat"), 0)
    area, pos = _numbers(lines[0]), _numbers(lines[2])
    data = np.array([_numbers(ln, drop_label=False) for ln in lines[5:] if ln.strip()])
    return dict(time=data[:, 0], current=data[:, 1:], pos=pos, area=area, steady=data[:, 1:].mean(axis=0))


def read_population(directory: str) -> Dict[str, np.ndarray]:
    """``population_profile.dat`` (monte_carlo_results.py:70-109): line 0 areas, 2 dy, 4 slab centres, 6 column names."""
    lines = _rows(os.path.join(directory, "population_profile.dat"), 0)
    area, dy, pos = _numbers(lines[0]), _numbers(lines[2]), _numbers(lines[4])
    data = np.array([_numbers(ln, drop_label=False) for ln in lines[7:] if ln.strip()])
    return dict(time=data[:, 0], pop=data[:, 1:], area=area, dy=dy, pos=pos, avg_pop=data[:, 1:].mean(axis=0))


def read_scatterer_stats(directory: str) -> Dict[str, np.ndarray]:
    """``scatterer_statistics.dat`` (monte_carlo_results.py:111-126): csv with a header row."""
    lines = _rows(os.path.join(directory, "scatterer_statistics.dat"), 0)
    names = lines[0].split(",")
    data = np.array([_numbers(ln, drop_label=False) for ln in lines[1:] if ln.strip()])
    return {n: data[:, i] for i, n in enumerate(names)}


def contact_diffusion(directory: str) -> Dict[str, np.ndarray]:
    """Fick's law on the steady state of a contact run (monte_carlo_results.py:428-437): the time-averaged current
    density through interface i divided by the density gradient between slabs i and i+1.  (The reference divides without
    the minus sign of Fick's law; so does this function, to give the numbers its script gives.)"""
    c, p = read_current(directory), read_population(directory)
    gradient = np.diff(p["avg_pop"]) / p["dy"][:-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        d = c["steady"] / gradient
    return dict(pos=c["pos"], diffusion=d, steady_current=c["steady"], gradient=gradient)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("directory")
    ap.add_argument("--kubo", action="store_true", help="diffusion coefficients from particle_dispalcement.avg.squared.dat")
    ap.add_argument("--diffusion", action="store_true", help="Fick's-law coefficients from a contact run")
    ap.add_argument("--skip", type=float, default=0.0, help="fraction of the run ignored by the --kubo fit")
    a = ap.parse_args(argv)
    out = {}
    if a.kubo:
        k = kubo_diffusion(a.directory, a.skip)
        out["kubo"] = {"D_m2_per_s": {ax: float(v) for ax, v in k["D"].items()},
                       "msd_over_time_last": {ax: float(v[-1]) for ax, v in k["msd_over_time"].items()},
                       "rows": int(len(k["time"]))}
    if a.diffusion:
        d = contact_diffusion(a.directory)
        out["contacts"] = {k: [float(x) for x in v] for k, v in d.items()}
    print(json.dumps(out, indent=1))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
