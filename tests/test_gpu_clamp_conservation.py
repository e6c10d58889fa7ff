"""GPU parity in the entropy-clamp regime (EquationsOfState.jl:152-155), the conservation budget of the
two-phase model ON THE GPU RESULT (north_star: "mass/momentum/energy conservation checked"), and update_cell (main.jl:30-60)
through the Python mirror."""
import numpy as np
import pytest

from util import random_mph_prims, relerr

pytestmark = pytest.mark.gpu
TOL_OP, TOL_RUN = 1e-12, 1e-9


def _cold_states(oracle, oe, rng, n, both_phases=True):
    """admissible two-phase states whose internal energy is pushed below the cold curve -> S' clamps at 1e-6"""
    P = random_mph_prims(rng, n)
    P[:, 5] = rng.uniform(0, 4e-4, n); P[:, 20] = P[:, 5]
    Q, _ = oracle.prim2cons(oe, oracle.MPH30, P)
    Q = Q.copy()
    shift = rng.uniform(0.4, 2.0, n)
    Q[:, 5] -= Q[:, 1] * shift
    if both_phases:
        Q[:, 20] -= Q[:, 16] * shift
    return Q


def test_clamped_states_cell_functions(gpu, oracle):
    hs = gpu
    rng = np.random.default_rng(21)
    eos = (hs.Barton2009(), hs.Barton2009())
    oe = [oracle.barton2009(), oracle.barton2009()]
    for both in (True, False):      # False: phase 1 clamped, phase 2 hot -> the interface-stress weights T1, T2 differ by 1e6
        Q = _cold_states(oracle, oe, rng, 300, both)
        Po, st = oracle.cons2prim(oe, oracle.MPH30, Q)
        assert st == 0
        S_clamp = oe[0][2] * np.log(1e-6)
        assert np.allclose(Po[:, 5], S_clamp, rtol=1e-13)            # the oracle is in the clamp branch
        P = hs.cons2prim_mph(eos, Q)
        assert relerr(P, Po) < TOL_OP                                 # S = cv log(1e-6): no amplification here
        assert np.allclose(P[:, 5], S_clamp, rtol=1e-15)
        F = hs.flux_mph(eos, Q); Fo, _ = oracle.flux(oe, oracle.MPH30, Q)
        assert relerr(F, Fo) < TOL_OP
        col = hs.noncons_flux(eos, Q, dense=False); colo, _ = oracle.noncons_cols(oe, Q)
        assert relerr(col, colo) < TOL_OP                             # T enters here (HyperelasticityMPh.jl:212-217)
        eg = hs.get_eigvals(eos, Q); ego, _ = oracle.get_eigvals(oe, oracle.MPH30, Q)
        assert relerr(eg, ego, per_var=False) < TOL_OP
        # one path-conservative HLL face between neighbouring clamped states (18 quadrature states, all clamped)
        Ql, Qr = Q[:-1], Q[1:]
        el, er = ego[:-1], ego[1:]
        _, dm, dp, s = hs.hll(eos, Ql, Qr, [el, er], return_speeds=True)
        _, dmo, dpo, so, st = oracle.hll(oe, Ql, Qr, el, er)
        assert relerr(s, so, per_var=False) < TOL_OP
        assert relerr(dm, dmo) < TOL_OP and relerr(dp, dpo) < TOL_OP


def test_two_phase_run_through_clamped_states(gpu, oracle):
    """10 steps of the fused two-phase kernel on a cold Riemann problem (left state clamped in phase 1, right state clamped in
    both phases): <= 1e-12 after one step, <= 1e-9 after 10."""
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    oe = [oracle.barton2009(), oracle.barton2009()]
    Ql, Qr = hs.initial_states(eos, 6)
    Ql = Ql.copy(); Qr = Qr.copy()
    Ql[5] -= 2.0 * Ql[1]
    Qr[5] -= 0.5 * Qr[1]; Qr[20] -= 0.5 * Qr[16]
    nx = 256
    Q0 = hs.initial_condition(Ql, Qr, nx)
    P0, _ = oracle.cons2prim(oe, oracle.MPH30, Q0)
    S_clamp = oe[0][2] * np.log(1e-6)
    assert np.isclose(P0[0, 5], S_clamp) and not np.isclose(P0[0, 20], S_clamp) and np.isclose(P0[-1, 5], S_clamp) and np.isclose(P0[-1, 20], S_clamp)
    for flux, fk in (("hll", oracle.HLL), ("lxf", oracle.LXF)):
        one = oracle.run(oe, oracle.MPH30, fk, Q0, 0.6, 1.0 / nx, 1e9, 1, nthreads=8)
        ten = oracle.run(oe, oracle.MPH30, fk, Q0, 0.6, 1.0 / nx, 1e9, 10, nthreads=8)
        assert one["status"] == 0 and ten["status"] == 0
        with hs.Solver(eos, nx) as sol:
            sol.upload(Q0)
            dt1 = sol.step(flux, 0.6, 1.0 / nx)
            assert abs(dt1[0] - one["dt"][0, 0]) <= 1e-13 * dt1[0]
            assert relerr(sol.download(), one["Q"]) < TOL_OP
            sol.upload(Q0)
            sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=10)
            assert relerr(sol.download(), ten["Q"]) < TOL_RUN


@pytest.mark.parametrize("flux", ["hll", "lxf"])
def test_two_phase_conservation_budget_on_gpu(gpu, flux):
    """per-phase mass (Q[2], Q[17]) is conserved, mixture momentum (Q[3:5]+Q[18:20]) and mixture energy
    (Q[6]+Q[21]) of the interior change exactly by the boundary-flux budget -(t/dx)(F_R - F_L); asserted on the CUDA result at
    nx = 10^4 (the interior waves never reach the frozen boundary cells in 60 steps)."""
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    Ql, Qr = hs.initial_states(eos, 6)
    nx, nsteps = 10000, 60
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0)
        sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps)
        Q = sol.download(); t = float(sol.t[0])
    Fl = hs.flux_mph(eos, Ql); Fr = hs.flux_mph(eos, Qr)
    interior = slice(1, nx - 1)
    import math
    for cols in ([1], [16], [2, 17], [3, 18], [4, 19], [5, 20]):
        tot0 = math.fsum(Q0[interior][:, cols].ravel()); tot1 = math.fsum(Q[interior][:, cols].ravel())
        budget = -(t * nx) * (Fr[cols].sum() - Fl[cols].sum())
        assert abs((tot1 - tot0) - budget) < 1e-10 * max(1.0, abs(tot0)), (cols, tot1 - tot0, budget)
    # volume fractions stay a partition of unity; rho F components are NOT conserved (non-conservative coupling): not asserted
    assert np.abs(Q[:, 0] + Q[:, 15] - 1.0).max() < 1e-13


def test_update_cell_mirror_both_methods(gpu, oracle):
    """main.jl:30-41 (LxF) and :43-60 (HLL) through hyperelasticsolver_b200.update_cell on 3-cell stencils, against the
    oracle's hll / lxf composed the same way, and against one fused step of the solver (cell 2 of a 3... 5-cell grid)."""
    hs = gpu
    rng = np.random.default_rng(5)
    eos = (hs.Barton2009(), hs.Barton2009())
    oe = [oracle.barton2009(), oracle.barton2009()]
    P = random_mph_prims(rng, 3, spread=0.03)
    Q3, _ = oracle.prim2cons(oe, oracle.MPH30, P)
    eig = hs.get_eigvals(eos, Q3)
    dtdx = 0.02
    # HLL method
    qn = hs.update_cell(Q3, hs.hll, eig, dtdx, eos)
    ego, _ = oracle.get_eigvals(oe, oracle.MPH30, Q3)
    cl, dml, dpl, _, _ = oracle.hll(oe, Q3[0:1], Q3[1:2], ego[0:1], ego[1:2])
    cr, dmr, dpr, _, _ = oracle.hll(oe, Q3[1:2], Q3[2:3], ego[1:2], ego[2:3])
    ref = Q3[1] - dtdx * ((cr[0] - cl[0]) + (dmr[0] + dpl[0]))
    assert relerr(qn[None, :], ref[None, :]) < TOL_OP
    # LxF method
    lam = 1.0 / dtdx
    qn = hs.update_cell(Q3, hs.lxf, lam, eos)
    cl, dml, dpl, _ = oracle.lxf(oe, Q3[0:1], Q3[1:2], lam)
    cr, dmr, dpr, _ = oracle.lxf(oe, Q3[1:2], Q3[2:3], lam)
    ref = Q3[1] - 1.0 / lam * ((cr[0] - cl[0]) + (dmr[0] + dpl[0]))
    assert relerr(qn[None, :], ref[None, :]) < TOL_OP
    # the fused step computes the same thing: 3-cell grid, middle cell, dt/dx from the step's own dt
    with hs.Solver(eos, 3) as sol:
        sol.upload(Q3)
        dt = sol.step("hll", 0.6, 1.0)
        Qs = sol.download()
    qn = hs.update_cell(Q3, hs.hll, eig, dt[0] / 1.0, eos)
    assert relerr(Qs[1:2], qn[None, :]) < TOL_OP
    assert np.array_equal(Qs[0], Q3[0]) and np.array_equal(Qs[2], Q3[2])     # main.jl:219-220
