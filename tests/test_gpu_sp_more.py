"""GPU: more single-phase coverage -- tile-boundary sizes x {HLL, LxF} x {default, generic EoS exponents},
and a single-phase ensemble whose problems reach t_end at different step counts."""
import numpy as np
import pytest

from util import random_sp_prims, relerr

pytestmark = pytest.mark.gpu


def test_single_phase_matrix(gpu, oracle):
    hs = gpu
    checks = 0
    for kind in ("default", "generic"):
        eos = hs.Barton2009() if kind == "default" else hs.Barton2009(_c0=6.22, _cv=9.0e-4, _b0=3.16, _beta=3.577, _gamma=2.088)
        oe = [oracle.barton2009()] if kind == "default" else [oracle.barton2009(c0=6.22, cv=9.0e-4, b0=3.16, beta=3.577, gamma=2.088)]
        for flux, fk in (("hll", oracle.HLL), ("lxf", oracle.LXF)):
            for nx in (3, 127, 128, 129, 300, 1000):
                Ql, Qr = hs.hyperelasticity.initial_states(eos, 2)
                Q0 = hs.initial_condition(Ql, Qr, nx)
                steps = 6
                ref = oracle.run(oe, oracle.SP13, fk, Q0, 0.6, 1.0 / nx, 1e9, steps, nthreads=8)
                with hs.Solver(eos, nx, model=hs.SP13) as sol:
                    sol.upload(Q0)
                    hist = sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=steps, record_dt=True)
                    Q = sol.download()
                assert np.allclose(hist[0], ref["dt"][0], rtol=1e-12, atol=0), (kind, flux, nx)
                assert relerr(Q, ref["Q"]) < 1e-10, (kind, flux, nx, relerr(Q, ref["Q"]))
                assert np.array_equal(Q[0], Q0[0]) and np.array_equal(Q[-1], Q0[-1])
                checks += 1
    # ensemble with per-problem dt and different finishing steps
    rng = np.random.default_rng(3)
    eos = hs.Barton2009(); oe = [oracle.barton2009()]
    nprob, nx = 7, 260
    Ql = hs.hyperelasticity.prim2cons(eos, random_sp_prims(rng, nprob, spread=0.03)); Qr = hs.hyperelasticity.prim2cons(eos, random_sp_prims(rng, nprob, spread=0.03))
    Q0 = np.stack([hs.initial_condition(Ql[i], Qr[i], nx) for i in range(nprob)])
    t_end = 0.004
    ref = oracle.run(oe, oracle.SP13, oracle.HLL, Q0, 0.6, 1.0 / nx, t_end, 500, nthreads=8)
    assert ref["status"] == 0 and len(set(ref["steps"].tolist())) > 1
    with hs.Solver(eos, nx, nprob=nprob, model=hs.SP13) as sol:
        sol.upload(Q0)
        sol.advance(t_end, "hll", 0.6, 1.0 / nx, max_steps=500)
        Q = sol.download()
        assert np.array_equal(sol.steps, ref["steps"]) and np.allclose(sol.t, ref["t"], rtol=1e-11)
    assert relerr(Q.reshape(-1, 13), ref["Q"].reshape(-1, 13)) < 1e-9
    checks += 1
    assert checks == 25
