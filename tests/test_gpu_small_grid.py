"""GPU: the device-resident loop on grids too small to fill the GPU.
  * k_step_qp (quadrature-parallel two-phase step, one warp per quadrature node) must be BIT-IDENTICAL to the fused kernel k_step
    -- same fused multiply-adds in the same order -- so that the size-based switch inside hsd_step is invisible;
  * hs_advance replaying a six-step CUDA graph must be bit-identical to launching the steps one by one, keep the
    `while t < T` semantics (no clipping of the last step, main.jl:202,214) and the dt history;
  * the run main.jl ships (config 0: 641 steps to t = 0.06003958139325407)."""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from util import random_mph_prims, relerr

pytestmark = pytest.mark.gpu


@contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    os.environ.update({k: str(v) for k, v in kw.items()})
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _eos(hs, kind):
    if kind == "default":
        return (hs.Barton2009(), hs.Barton2009())
    return (hs.Barton2009(), hs.Barton2009(_rho0=8.93, _c0=6.22, _cv=9.0e-4, _t0=300, _b0=3.16, _alpha=1, _beta=3.577, _gamma=2.088))


def _smooth_two_phase(hs, eos, nx, tc=6):
    Ql, Qr = hs.initial_states(eos, tc)
    x = (np.arange(nx) + 0.5) / nx
    w = (0.5 * (1 + np.tanh((x - 0.5) / 0.08)))[:, None]
    Pl, Pr = hs.cons2prim_mph(eos, Ql), hs.cons2prim_mph(eos, Qr)
    P = (1 - w) * Pl[None, :] + w * Pr[None, :]
    P[:, 15] = 1.0 - P[:, 0]
    P[:, 2] += 0.3 * np.sin(6 * np.pi * x); P[:, 17] = P[:, 2]
    return hs.prim2cons_mph(eos, P)


def _run(hs, eos, Q0, nsteps, flux, nprob=1, **envkw):
    nx = Q0.shape[-2]
    with env(**envkw), hs.Solver(eos, nx, nprob=nprob) as sol:
        sol.upload(Q0)
        hist = sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps, record_dt=True)
        return sol.download(), hist, sol.t.copy(), sol.steps.copy()


@pytest.mark.parametrize("kind", ["default", "hetero"])
@pytest.mark.parametrize("flux", ["hll", "lxf"])
@pytest.mark.parametrize("nx", [3, 16, 17, 30, 257, 1000, 2048])
def test_quadrature_parallel_kernel_bit_identical_to_fused_kernel(gpu, kind, flux, nx):
    hs = gpu
    eos = _eos(hs, kind)
    Q0 = _smooth_two_phase(hs, eos, nx, 7 if kind == "hetero" else 6)
    nsteps = 9
    Qa, ha, ta, sa = _run(hs, eos, Q0, nsteps, flux, HS_QP_MAX_CELLS=0, HS_GRAPH=0)          # fused kernel k_step
    for loop in (0, 1):   # k_step_qp launched per step / k_step_qp_loop: all steps in one cooperative launch with grid barriers
        Qb, hb, tb, sb = _run(hs, eos, Q0, nsteps, flux, HS_QP_MAX_CELLS=1 << 20, HS_GRAPH=0, HS_QP_LOOP=loop)
        assert np.array_equal(ha, hb), (loop, ha, hb)
        assert np.array_equal(Qa, Qb), (loop, np.abs(Qa - Qb).max())
        assert np.array_equal(ta, tb) and np.array_equal(sa, sb)


def test_quadrature_parallel_kernel_ensemble_and_slabs(gpu):
    """several small problems with their own dt, finishing at different steps; and the ghost-cell windows of a slab decomposition"""
    hs = gpu
    from hyperelasticsolver_b200 import _lib as L
    from hyperelasticsolver_b200.slab import CudaKernels
    from tools.two_slabs_one_device import run_slabs
    eos = _eos(hs, "default")
    rng = np.random.default_rng(3)
    nprob, nx = 6, 130
    P = random_mph_prims(rng, 2 * nprob, spread=0.03).reshape(nprob, 2, 30)
    Qlr = hs.prim2cons_mph(eos, P)
    Q0 = np.where((np.arange(nx) < nx / 2)[None, :, None], Qlr[:, None, 0, :], Qlr[:, None, 1, :]).copy()
    outs = []
    for cap, loop in ((0, 0), (1 << 20, 0), (1 << 20, 1)):
        with env(HS_QP_MAX_CELLS=cap, HS_QP_LOOP=loop), hs.Solver(eos, nx, nprob=nprob) as sol:
            sol.upload(Q0)
            lam = sol.wave_speeds()
            t_end = 7.3 * 0.6 * (1.0 / nx) / lam.max()
            hist = sol.advance(t_end, "hll", 0.6, 1.0 / nx, max_steps=60, record_dt=True)
            outs.append((sol.download(), hist, sol.t.copy(), sol.steps.copy()))
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert np.array_equal(x, y)
    assert len(set(outs[0][3].tolist())) > 1          # the problems really stop at different step counts
    # slab windows with ghost cells through hsd_step (every slab small enough for k_step_qp), against the fused kernel on the whole grid
    nx = 1500
    Q0 = _smooth_two_phase(hs, eos, nx)
    ref, *_ = _run(hs, eos, Q0, 7, "hll", HS_QP_MAX_CELLS=0)
    kern = CudaKernels(eos, hs.MPH30, "cuda:0")
    with env(HS_QP_MAX_CELLS=1 << 20):
        Q, _, nlocs = run_slabs(kern, Q0, 3, 7, L.HLL)
    assert np.array_equal(Q, ref)


@pytest.mark.parametrize("model,nx", [("mph30", 1000), ("mph30", 5000), ("sp13", 1000), ("sp13", 70000)])
def test_graph_replay_bit_identical_and_termination(gpu, model, nx):
    hs = gpu
    if model == "mph30":
        eos = _eos(hs, "default"); Q0 = _smooth_two_phase(hs, eos, nx); hm = hs.MPH30
    else:
        eos = hs.Barton2009(); Ql, Qr = hs.hyperelasticity.initial_states(eos, 1); Q0 = hs.initial_condition(Ql, Qr, nx); hm = hs.SP13
    res = []
    for graph, loop in ((0, 0), (1, 0), (1, 1)):
        with env(HS_GRAPH=graph, HS_QP_LOOP=loop), hs.Solver(eos, nx, model=hm) as sol:
            sol.upload(Q0)
            lam = sol.wave_speeds()[0]
            t_end = 100.4 * 0.6 * (1.0 / nx) / lam            # ~100-140 steps; the clock decides, not max_steps
            hist = sol.advance(t_end, "hll", 0.6, 1.0 / nx, max_steps=400, record_dt=True)
            n = int(sol.steps[0])
            assert 90 <= n <= 200
            assert sol.t[0] >= t_end and sol.t[0] - hist[0, n - 1] < t_end      # last step not clipped, none taken after t >= T
            assert np.all(hist[0, n:] == 0.0) and np.all(hist[0, :n] > 0.0)
            assert abs(hist[0, :n].sum() - sol.t[0]) < 1e-12 * sol.t[0]
            # a second stretch from a non-zero clock and an odd position in the buffer rotation
            sol.advance(2 * t_end, "hll", 0.6, 1.0 / nx, max_steps=37)
            res.append((sol.download(), hist.copy(), float(sol.t[0]), int(sol.steps[0])))
    for r in res[1:]:
        assert np.array_equal(res[0][0], r[0]) and np.array_equal(res[0][1], r[1]) and res[0][2:] == r[2:]


def test_config0_through_graph_and_small_grid_kernel(gpu, oracle):
    """BASELINE config 0 as the default build runs it (k_step_qp replayed from a CUDA graph): SURVEY.md B.6 golden values and the oracle"""
    hs = gpu
    eos = _eos(hs, "default")
    Ql, Qr = hs.initial_states(eos, 6)
    nx = 1000
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0)
        hist = sol.advance(0.06, "hll", 0.6, 1.0 / nx, max_steps=1000, record_dt=True)
        Q = sol.download()
        assert int(sol.steps[0]) == 641
        assert abs(sol.t[0] - 0.06003958139325407) < 1e-13
        assert abs(hist[0, 640] - 9.427169653771197e-5) < 1e-15
    ref = [0.22565851708953968, 1.9195677660678498, 0.6660210927551466, 0.3926551441509054, 0.855227842806319, 1.7104174953968945]
    assert np.allclose(Q[500, :6], ref, rtol=1e-10, atol=0)
    r = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, 0.06, 1000, nthreads=oracle.hardware_threads())
    assert relerr(Q, r["Q"]) < 1e-9
