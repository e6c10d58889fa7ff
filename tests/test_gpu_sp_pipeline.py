"""GPU: the single-phase TMA tile pipeline (k_step_sp) against the plain fused kernel (k_step<SP13>, HS_SP_TMA=0).

Both run the same inline arithmetic, so the results must agree to the last bit or two (the compiler may contract
a multiply-add differently in the two kernels); the plain kernel is itself checked against the oracle elsewhere, and
every other single-phase GPU test in this directory runs through the pipeline by default.  Covered here: both
copy flavours (tensor-map tiles where the stride is even, row copies otherwise or with HS_SP_TMA2D=0), every
alignment case of the row copies (odd / even ncells, odd / even problem offsets), grids smaller than a tile, tiles
whose copy window would cross the end of the arrays (loaded by the threads), 1 / 3 / 8 tiles per block, both
fluxes, generic EoS exponents, ensembles whose problems stop at different steps, slab ghost cells.
"""
import os

import numpy as np
import pytest

from util import random_sp_prims, relerr

pytestmark = pytest.mark.gpu


def _run(hs, eos, Q0, nx, nprob, flux, steps, t_end=1e9, tma="1", tiles=None, tma2d=None):
    old = {k: os.environ.get(k) for k in ("HS_SP_TMA", "HS_SP_TILES", "HS_SP_TMA2D")}
    os.environ["HS_SP_TMA"] = tma
    if tma2d is None:
        os.environ.pop("HS_SP_TMA2D", None)
    else:
        os.environ["HS_SP_TMA2D"] = tma2d
    if tiles is None:
        os.environ.pop("HS_SP_TILES", None)
    else:
        os.environ["HS_SP_TILES"] = str(tiles)
    try:
        with hs.Solver(eos, nx, nprob=nprob, model=hs.SP13) as sol:
            sol.upload(Q0)
            hist = sol.advance(t_end, flux, 0.6, 1.0 / nx, max_steps=steps, record_dt=True)
            return sol.download(), hist, sol.steps.copy(), sol.t.copy()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _close(a, b, tol=4e-15):
    return relerr(a, b) <= tol


@pytest.mark.parametrize("flux", ["hll", "lxf"])
def test_pipeline_matches_plain_kernel_single_grid(gpu, flux):
    hs = gpu
    eos = hs.Barton2009()
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 2)
    for nx in (3, 4, 127, 128, 129, 130, 253, 254, 255, 379, 380, 1000, 5003, 40000):
        Q0 = hs.initial_condition(Ql, Qr, nx)
        ref, href, _, _ = _run(hs, eos, Q0, nx, 1, flux, 7, tma="0")
        for tiles, tma2d in ((1, None), (3, None), (None, None), (3, "0"), (None, "0")):   # tensor-map tiles (even nx) / row copies
            Q, h, _, _ = _run(hs, eos, Q0, nx, 1, flux, 7, tiles=tiles, tma2d=tma2d)
            assert np.allclose(h, href, rtol=1e-14, atol=0), (nx, tiles, tma2d)
            assert _close(Q, ref), (nx, tiles, tma2d, relerr(Q, ref))
            assert np.array_equal(Q[0], Q0[0]) and np.array_equal(Q[-1], Q0[-1])   # frozen boundary cells


def test_pipeline_large_grid_many_tiles_per_block(gpu):
    hs = gpu
    eos = hs.Barton2009()
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
    nx = (1 << 21) + 77          # > 2 * 4 * 148 * 8 tiles: the default 8 tiles per block, odd cell count
    Q0 = hs.initial_condition(Ql, Qr, nx)
    # a smooth perturbation so that every tile differs
    x = np.linspace(0, 1, nx)[:, None]
    Q0 = Q0 * (1.0 + 0.01 * np.sin(40 * np.pi * x))
    ref, href, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 5, tma="0")
    for tiles in (None, 16, 5):
        Q, h, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 5, tiles=tiles)
        assert np.allclose(h, href, rtol=1e-14, atol=0)
        assert _close(Q, ref), relerr(Q, ref)
    # even cell count: the tensor-map tile copies, against the same plain kernel
    nx = (1 << 21) + 78
    Q0 = hs.initial_condition(Ql, Qr, nx) * (1.0 + 0.01 * np.sin(40 * np.pi * np.linspace(0, 1, nx)[:, None]))
    ref, href, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 5, tma="0")
    for tma2d in (None, "0"):
        Q, h, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 5, tma2d=tma2d)
        assert np.allclose(h, href, rtol=1e-14, atol=0)
        assert _close(Q, ref), (tma2d, relerr(Q, ref))


def test_pipeline_generic_exponents(gpu):
    hs = gpu
    eos = hs.Barton2009(_c0=6.22, _cv=9.0e-4, _b0=3.16, _beta=3.577, _gamma=2.088)
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 2)
    nx = 3001
    Q0 = hs.initial_condition(Ql, Qr, nx)
    ref, href, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 6, tma="0")
    Q, h, _, _ = _run(hs, eos, Q0, nx, 1, "hll", 6, tiles=4)
    assert np.allclose(h, href, rtol=1e-14, atol=0) and _close(Q, ref, 1e-14)


# (301, 8) / (7, 300): ODD problem length with an EVEN total -- tiles start on odd columns there, which the tensor-map copies
# cannot do (16-byte alignment of the box start): the launcher must pick the row copies (an illegal-instruction fault before round 2)
@pytest.mark.parametrize("nx,nprob", [(301, 9), (300, 9), (4096, 40), (125, 33), (3, 5), (301, 8), (7, 300)])
def test_pipeline_ensembles_with_different_stopping_steps(gpu, nx, nprob):
    hs = gpu
    eos = hs.Barton2009()
    rng = np.random.default_rng(nx + nprob)
    Ql = hs.hyperelasticity.prim2cons(eos, random_sp_prims(rng, nprob, spread=0.03))
    Qr = hs.hyperelasticity.prim2cons(eos, random_sp_prims(rng, nprob, spread=0.03))
    Q0 = np.stack([hs.initial_condition(Ql[i], Qr[i], nx) for i in range(nprob)])
    t_end = 12 * 0.6 / nx / 6.0       # about a dozen steps; problems stop at different counts
    ref, href, sref, tref = _run(hs, eos, Q0, nx, nprob, "hll", 60, t_end=t_end, tma="0")
    assert len(set(sref.tolist())) > 1 or nprob < 4
    for tiles, tma2d in ((1, None), (2, None), (None, None), (None, "0")):
        Q, h, s, t = _run(hs, eos, Q0, nx, nprob, "hll", 60, t_end=t_end, tiles=tiles, tma2d=tma2d)
        assert np.array_equal(s, sref), (tiles, s, sref)
        assert np.allclose(t, tref, rtol=1e-13, atol=0)
        assert np.allclose(h, href, rtol=1e-13, atol=0)
        assert _close(Q.reshape(-1, 13), ref.reshape(-1, 13), 1e-13), relerr(Q.reshape(-1, 13), ref.reshape(-1, 13))
