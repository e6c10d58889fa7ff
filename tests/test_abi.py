"""CPU: the C-ABI library loads and exports every symbol include/hyperelastic_b200.h declares;
argument errors are reported; without a GPU compute calls fail loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hyperelastic_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hs(?:d|2d)?_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(hs):
    from hyperelasticsolver_b200 import _lib as L
    names = _declared_symbols()
    assert len(names) >= 25
    raw = C.CDLL(L.SO_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in L.SIGNATURES, f"{n} has no ctypes signature"
    assert set(L.SIGNATURES) == set(names)


def test_struct_layouts(hs):
    from hyperelasticsolver_b200 import _lib as L
    assert C.sizeof(L.Barton2009) == 80
    e = L.Barton2009()
    assert e.as_tuple() == (8.93, 4.6, 3.9e-4, 300.0, 2.1, 1.0, 3.0, 2.0, 2.1 ** 2, 4.6 ** 2 - (4 / 3) * 2.1 ** 2)
    assert C.sizeof(L.HsdProblem) == 4 * 4 + 3 * 8 + 2 * 20 * 8


def test_cache_row_count_matches_header(hs):
    """hsd_naux() (the layout the library was built with) == HS_NAUX of the header at its default build option;
    the Hank2016 block is 5 doubles in the reference's field order."""
    from hyperelasticsolver_b200 import _lib as L
    txt = open(os.path.join(ROOT, "include", "hyperelastic_b200.h")).read()
    crow = int(re.search(r"#ifndef HS_SP_CROW\s*\n#define HS_SP_CROW (\d)", txt).group(1))
    assert "#define HS_NAUX_SP (HS_SP_CROW ? 5 : 6)" in txt
    lib = L.lib()
    assert lib.hsd_naux(L.SP13) == (5 if crow else 6) == L.NAUX[L.SP13]
    assert lib.hsd_naux(L.MPH30) == 2 == L.NAUX[L.MPH30]
    assert C.sizeof(L.Hank2016) == 40
    h = L.Hank2016()
    assert (h.rho0, h.mu, h.gamma, h.pres_inf, h.a) == (2.7, 26e9, 3.4, 21.5e9, 0.5)   # EquationsOfState.jl:312-318


def test_problem_init_and_arg_errors(hs):
    from hyperelasticsolver_b200 import _lib as L
    lib = L.lib()
    p = L.HsdProblem()
    eos2 = L.eos_array((hs.Barton2009(), hs.Barton2009()), L.MPH30)
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, eos2, 2, 1000, 1) == 0
    assert (p.model, p.nphase, p.gen, p.ncells, p.nprob, p.stride) == (1, 2, 0, 1000, 1, 1000)
    het = L.eos_array((hs.Barton2009(), hs.Barton2009(_beta=3.577, _gamma=2.088)), L.MPH30)
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, het, 2, 1000, 4) == 0 and p.gen == 1 and p.stride == 4000
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, eos2, 1, 1000, 1) == L.HS_ERR_ARG
    assert lib.hsd_problem_init(C.byref(p), 7, eos2, 2, 1000, 1) == L.HS_ERR_ARG
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, eos2, 2, 2, 1) == L.HS_ERR_ARG
    assert b"ncells" in lib.hs_last_error()
    # the kernels index cells with 32 bits: 65 536 x 4 096 = 2^28 is fine, 2^31 cells are not (more than one device holds anyway)
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, eos2, 2, 4096, 65536) == 0 and p.stride == 1 << 28
    assert lib.hsd_problem_init(C.byref(p), L.MPH30, eos2, 2, 1 << 16, 1 << 15) == L.HS_ERR_ARG
    with pytest.raises(ValueError):
        L.eos_array((hs.Barton2009(),), L.MPH30)


def test_no_cpu_fallback(hs):
    """On a machine without a GPU every compute entry point must fail with HS_ERR_CUDA."""
    from hyperelasticsolver_b200 import _lib as L
    if L.lib().hs_device_count() > 0:
        pytest.skip("GPU present")
    eos = (hs.Barton2009(), hs.Barton2009())
    with pytest.raises(hs.HyperelasticError) as ei:
        hs.cons2prim_mph(eos, np.ones(30))
    assert ei.value.code == L.HS_ERR_CUDA
    with pytest.raises(hs.HyperelasticError):
        hs.Solver(eos, 100)
    with pytest.raises(hs.HyperelasticError) as ei:
        hs.Solver2D(eos, 16, 16)
    assert ei.value.code == L.HS_ERR_CUDA
    with pytest.raises(hs.HyperelasticError) as ei:
        hs.register_host(np.zeros(64))
    assert ei.value.code == L.HS_ERR_CUDA
    # argument errors are reported before any device work
    assert L.lib().hs_step_host(None, L.HLL, 0.6, 0.1, None, None, None) == L.HS_ERR_ARG
    assert L.lib().hs2d_step(None, L.HLL, 0.6, 0.1, 0.1, None) == L.HS_ERR_ARG


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing in the product package may reference it."""
    pkg = os.path.join(ROOT, "hyperelasticsolver_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src and "oracle/" not in src.replace("dual-number oracle", ""), f
