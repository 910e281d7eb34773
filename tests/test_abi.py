"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/cntmc.h declares, parses the
reference's input.json schema, reports errors like the reference does -- and refuses to compute without CUDA."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from cnt_film_monte_carlo_b200 import _lib, build
from cnt_film_monte_carlo_b200.engine import CntmcError, Engine
from conftest import ROOT, base_mc


@pytest.fixture(scope="module", autouse=True)
def built():
    build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cntmc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cntmc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH]).decode()
    exported = set(re.findall(r" T (cntmc_[a-z_0-9]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    lib = _lib.load()
    assert b"sm_100a" in lib.cntmc_version()


def test_library_holds_sm100a_code():
    out = subprocess.check_output(["/usr/local/cuda/bin/cuobjdump", "--list-elf", _lib.LIB_PATH]).decode()
    assert "sm_100a" in out


def test_create_accepts_whole_file_or_block_and_reads_the_schema():
    mc = base_mc()
    for cfg in ({"exciton monte carlo": mc, "cnts": {}}, mc):
        e = Engine(cfg)
        assert e.kubo_max_time() == 1e-9 and e.time_step() == 1e-13 and e.number_of_segments() == 10
        assert e.time() == 0.0 and e.number_of_particles() == 0 and e.hops() == 0
        e.close()


def test_error_reporting():
    with pytest.raises(CntmcError, match="exciton monte carlo") as ei:
        Engine({"something else": 1})  # main.cpp:47
    assert ei.value.code == -1
    with pytest.raises(CntmcError, match="json"):
        Engine("{ not json")
    mc = base_mc()
    del mc["max hopping radius [m]"]
    with pytest.raises(CntmcError, match="max hopping radius"):
        Engine(mc)
    with pytest.raises(CntmcError, match="significand is all ones"):   # the one kind of divisor div_by is not exact for
        Engine(base_mc(**{"exciton velocity [m/s]": float(np.nextafter(262144.0, 0.0))}))
    e = Engine(base_mc(**{"rate type": "dexter"}))
    e.set_mesh(np.zeros((3, 2, 2)), np.ones((3, 2, 2)))
    with pytest.raises(CntmcError, match="rate type must be one of"):  # monte_carlo.cpp:59
        e.kubo_init()
    e = Engine(base_mc())
    with pytest.raises(CntmcError, match="no mesh"):
        e.kubo_init()
    with pytest.raises(CntmcError, match="kubo_init first"):
        e.kubo_create_particles(10)
    with pytest.raises(CntmcError, match="unknown option"):
        e.set_option("nope", 1)
    e.set_option("chunk_steps", 16)
    assert e.get_option("chunk_steps") == 16
    for name in ("top_entries", "runs", "dirs"):  # the shortcuts of the hop kernel are on by default and can be switched off
        assert e.get_option(name) == 1
        e.set_option(name, 0)
        assert e.get_option(name) == 0


def test_mesh_loader_errors(tmp_path):
    e = Engine(base_mc(**{"mesh input directory": str(tmp_path / "missing")}))
    with pytest.raises(CntmcError, match="cannot open mesh file"):
        e.load_mesh()


def test_no_cpu_fallback():
    """Without a CUDA device the engine must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cnt_film_monte_carlo_b200 import film
    e = Engine(base_mc())
    e.set_mesh(*film.film(NT=4, NP=10, a=5.0, LX=30.0, LY=20.0, seed=1))
    with pytest.raises(CntmcError) as ei:
        e.kubo_init()
    assert ei.value.code == -2 and "no CPU execution path" in str(ei.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "cnt_film_monte_carlo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle|liboracle|libt0", text, re.M), f
