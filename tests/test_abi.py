"""CPU (-m "not gpu"): the C-ABI library loads, exports every symbol include/psph.h declares, its
structs have the layout the Python mirrors assume, and it refuses to run without a GPU (no CPU
fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import plume_sph_b200 as psph
from plume_sph_b200 import lib as plib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "psph.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(psph_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound():
    L = psph.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
        assert n in plib.API, "no ctypes prototype for " + n
    assert sorted(plib.API) == names


def test_struct_layouts():
    assert C.sizeof(psph.PsphConfig) == 384 or C.sizeof(psph.PsphConfig) % 8 == 0
    c = psph.config_default_3d()
    assert c.abi_version == 1 and c.dim == 3 and c.cutoff == 3.0 and c.cutoff2 == 8.0
    assert list(c.ncell) == [55, 55, 31] and c.delta == 800.0 and c.nranks == 1 and c.ptable_size == 4000000
    c1 = psph.config_default_1d()
    assert c1.dim == 1 and c1.use_gsph == 1 and c1.cutoff2 == 12.0 and c1.cfl == 0.1 and c1.ptable_size == 10000
    # the oracle's mirror of the option record has the same layout (the tests memmove between them)
    import oracle_py as op
    assert C.sizeof(op.Config) == C.sizeof(psph.PsphConfig)
    o = op.config_3d()
    for f, _ in psph.PsphConfig._fields_:
        a, b = getattr(c, f), getattr(o, f)
        if hasattr(a, "__len__"):
            assert list(a) == list(b), f
        else:
            assert a == b, f
    assert psph.PARTICLE_DTYPE.itemsize == 328 and psph.CELL_ATTR_DTYPE.itemsize == 88


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(psph.PsphError):
        psph.Solver(psph.config_default_3d())


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "plume_sph_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), os.path.join(dirpath, f)


def test_header_compiles_as_c_and_cxx_with_the_mirrored_layouts(tmp_path):
    """include/psph.h is a plain C ABI: it compiles as C99 and as C++ on its own, and the compilers agree with the
    Python mirrors (ctypes / numpy) on the size of every record that crosses the boundary"""
    import subprocess
    import oracle_py as op
    sizes = {"psph_config": C.sizeof(psph.PsphConfig), "psph_particle": psph.PARTICLE_DTYPE.itemsize,
             "psph_cell_attr": psph.CELL_ATTR_DTYPE.itemsize, "psph_image": psph.IMAGE_DTYPE.itemsize,
             "psph_erupt": C.sizeof(psph.PsphErupt), "psph_growth": C.sizeof(psph.PsphGrowth)}
    assert C.sizeof(op.Erupt) == sizes["psph_erupt"] and C.sizeof(op.Growth) == sizes["psph_growth"]
    body = '#include "psph.h"\n' + "".join(
        "typedef char check_%s[sizeof(%s) == %d ? 1 : -1];\n" % (n, n, s) for n, s in sizes.items())
    body += "typedef char check_sph_row[PSPH_ROW_BYTES_SPH == %d ? 1 : -1];\n" % psph.PARTICLE_SPH_DTYPE.itemsize
    body += "int main(void) { return 0; }\n"
    for comp, ext, std in (("gcc", "c", "-std=c99"), ("g++", "cc", "-std=c++11")):
        src = tmp_path / ("abi_check." + ext)
        src.write_text(body)
        r = subprocess.run([comp, std, "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
