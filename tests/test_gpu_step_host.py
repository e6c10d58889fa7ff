"""GPU: the chunk-pipelined hs_step_host (H2D of chunk i+1 || kernels of chunk i || D2H of chunk i-1, speculative dt from the
max(lambda) the previous call produced) must be BIT-IDENTICAL to hs_upload + hs_step + hs_download -- on a confirmed hint,
on a refuted hint (state edited between calls) and on the first call -- and the windows it steps (even boundaries, ghost
cells at both ends) must cover every cell exactly once."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(hs, model, nx):
    if model == "mph30-hetero":      # different equations of state per phase, generic exponents (main.jl:133)
        eos = (hs.Barton2009(), hs.Barton2009(_rho0=8.93, _c0=6.22, _cv=9.0e-4, _t0=300, _b0=3.16, _alpha=1, _beta=3.577, _gamma=2.088))
        Ql, Qr = hs.initial_states(eos, 7); hm = hs.MPH30
    elif model == "mph30":
        eos = (hs.Barton2009(), hs.Barton2009()); Ql, Qr = hs.initial_states(eos, 6); hm = hs.MPH30
    else:
        eos = hs.Barton2009(); Ql, Qr = hs.hyperelasticity.initial_states(eos, 1); hm = hs.SP13
    x = (np.arange(nx) + 0.5) / nx
    w = (0.5 * (1 + np.tanh((x - 0.5) / 0.02)))[:, None]
    Q0 = (1 - w) * Ql[None, :] + w * Qr[None, :]           # every cell differs: a stale halo cell cannot hide
    return eos, hm, np.ascontiguousarray(Q0)


@pytest.fixture
def small_chunks():
    old = os.environ.get("HS_HOST_CHUNK")
    os.environ["HS_HOST_CHUNK"] = "2048"
    yield 2048
    if old is None:
        del os.environ["HS_HOST_CHUNK"]
    else:
        os.environ["HS_HOST_CHUNK"] = old


@pytest.mark.parametrize("model,nx,flux", [("sp13", 20000, "hll"), ("sp13", 6146, "lxf"), ("sp13", 4096, "hll"),
                                           ("mph30", 9000, "hll"), ("mph30", 4100, "lxf"), ("mph30-hetero", 5000, "hll")])
def test_pipelined_step_host_bit_identical(gpu, small_chunks, model, nx, flux):
    hs = gpu
    eos, hm, Q0 = _problem(hs, model, nx)
    nsteps = 6
    # reference: resident loop
    with hs.Solver(eos, nx, model=hm) as ref:
        ref.upload(Q0)
        dts = ref.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps, record_dt=True)[0]
        Qref = ref.download()
    with hs.Solver(eos, nx, model=hm) as sol:
        Qa = hs.register_host(Q0.copy()); Qb = hs.register_host(np.empty_like(Q0))
        for k in range(nsteps):
            _, dt = sol.step_host(Qa, Qb, flux, 0.6, 1.0 / nx)
            assert dt[0] == dts[k], (k, dt[0], dts[k])
            Qa, Qb = Qb, Qa
        calls, hits = sol.step_host_stats()
        assert calls == nsteps and hits == nsteps - 1            # only the first call has no hint
        assert np.array_equal(Qa, Qref), np.abs(Qa - Qref).max()
        # the context is left as upload + step would leave it: resident stepping continues from there
        sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=2)
        Qc = sol.download()
        hs.unregister_host(Qa); hs.unregister_host(Qb)
    with hs.Solver(eos, nx, model=hm) as ref:
        ref.upload(Q0); ref.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps + 2)
        assert np.array_equal(Qc, ref.download())


def test_pipelined_step_host_refuted_hint_and_aliasing(gpu, small_chunks):
    """the caller edits the state between two calls (so the hinted max(lambda) is wrong), and passes Qout = Qin"""
    hs = gpu
    nx = 12288
    eos, hm, Q0 = _problem(hs, "sp13", nx)
    with hs.Solver(eos, nx, model=hm) as sol, hs.Solver(eos, nx, model=hm) as ref:
        Q = Q0.copy()
        sol.step_host(Q, Q, "hll", 0.6, 1.0 / nx)                       # first call: no hint
        Q[nx // 3, 0] = 25.0 * Q[nx // 3, 3]                             # a fast cell (u1 ~ 25): max(lambda) of the uploaded data changes
        Qe = Q.copy()
        sol.step_host(Q, Q, "hll", 0.6, 1.0 / nx)                       # refuted -> redone on the device
        sol.step_host(Q, Q, "hll", 0.6, 1.0 / nx)                       # confirmed again
        calls, hits = sol.step_host_stats()
        assert (calls, hits) == (3, 1)
        ref.upload(Q0); ref.step("hll", 0.6, 1.0 / nx); Q1 = ref.download()
        Q1[nx // 3, 0] = 25.0 * Q1[nx // 3, 3]
        assert np.array_equal(Q1, Qe)
        ref.upload(Q1); ref.step("hll", 0.6, 1.0 / nx); ref.step("hll", 0.6, 1.0 / nx)
        assert np.array_equal(Q, ref.download())


def test_step_host_plain_paths(gpu, oracle):
    """odd cell counts, grids below two chunks and ensembles take upload + step + download; HS_HOST_PIPELINE=0 forces it"""
    hs = gpu
    for nx in (4097, 300):
        eos, hm, Q0 = _problem(hs, "sp13", nx)
        with hs.Solver(eos, nx, model=hm) as sol, hs.Solver(eos, nx, model=hm) as ref:
            Q1, dt = sol.step_host(Q0)
            assert sol.step_host_stats() == (0, 0)
            ref.upload(Q0); ref.step(); assert np.array_equal(Q1, ref.download())
    os.environ["HS_HOST_PIPELINE"] = "0"
    try:
        eos, hm, Q0 = _problem(hs, "sp13", 1 << 21)
        with hs.Solver(eos, 1 << 21, model=hm) as sol:
            Q1, _ = sol.step_host(Q0)
            assert sol.step_host_stats() == (0, 0)
    finally:
        del os.environ["HS_HOST_PIPELINE"]
    with hs.Solver(eos, 1 << 21, model=hm) as sol:       # default chunk size (2^21 cells: one chunk), pageable numpy arrays
        Q2, _ = sol.step_host(Q0)
        Q3, _ = sol.step_host(Q2)
        assert sol.step_host_stats() == (2, 1)
        assert np.array_equal(Q1, Q2)
    with hs.Solver(eos, 1 << 21, model=hm) as ref:
        ref.upload(Q0); ref.advance(1e9, "hll", 0.6, 1.0 / (1 << 21), max_steps=2)
        assert np.array_equal(Q3, ref.download())


@pytest.mark.parametrize("model,nx,nprob", [("sp13", 512, 24), ("sp13", 130, 40), ("mph30", 256, 20)])
def test_pipelined_step_host_ensemble(gpu, small_chunks, model, nx, nprob):
    """ensembles: groups of whole problems per chunk, per-problem hints; bit-identical to the resident loop, also when one
    problem of the ensemble is edited between two calls (every hint but one confirmed -> the step is redone)"""
    hs = gpu
    from util import random_mph_prims, random_sp_prims
    rng = np.random.default_rng(nx + nprob)
    if model == "mph30":
        eos = (hs.Barton2009(), hs.Barton2009()); hm = hs.MPH30
        Qlr = hs.prim2cons_mph(eos, random_mph_prims(rng, 2 * nprob, spread=0.03)).reshape(nprob, 2, 30)
    else:
        eos = hs.Barton2009(); hm = hs.SP13
        Qlr = hs.hyperelasticity.prim2cons(eos, random_sp_prims(rng, 2 * nprob, spread=0.03)).reshape(nprob, 2, 13)
    Q0 = np.where((np.arange(nx) < nx / 2)[None, :, None], Qlr[:, None, 0, :], Qlr[:, None, 1, :]).copy()
    nsteps = 5
    with hs.Solver(eos, nx, nprob=nprob, model=hm) as ref:
        ref.upload(Q0)
        dts = ref.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps, record_dt=True)
        Qref = ref.download()
    with hs.Solver(eos, nx, nprob=nprob, model=hm) as sol:
        Qa = hs.register_host(Q0.copy()); Qb = hs.register_host(np.empty_like(Q0))
        for k in range(nsteps):
            _, dt = sol.step_host(Qa, Qb, "hll", 0.6, 1.0 / nx)
            assert np.array_equal(dt, dts[:, k]), k
            Qa, Qb = Qb, Qa
        assert sol.step_host_stats() == (nsteps, nsteps - 1)
        assert np.array_equal(Qa, Qref)
        # edit one problem: its hint is refuted, the whole step is redone with the true values
        Qa[nprob // 2, :, 0 if model == "sp13" else 2] *= 1.5
        Qe = Qa.copy()
        sol.step_host(Qa, Qb, "hll", 0.6, 1.0 / nx)
        assert sol.step_host_stats() == (nsteps + 1, nsteps - 1)
        hs.unregister_host(Qa); hs.unregister_host(Qb)
    with hs.Solver(eos, nx, nprob=nprob, model=hm) as ref:
        ref.upload(Qe); ref.step("hll", 0.6, 1.0 / nx)
        assert np.array_equal(Qb, ref.download())
