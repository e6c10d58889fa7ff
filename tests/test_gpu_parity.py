"""GPU parity: the CUDA path (through the C ABI) against the dual-number CPU oracle on the same
seeded inputs.  Tolerances follow BASELINE.json north_star: <= 1e-12 relative per step,
<= 1e-9 after N steps, on conserved variables."""
import numpy as np
import pytest

from util import random_mph_prims, random_sp_prims, relerr

pytestmark = pytest.mark.gpu

TOL_OP = 1e-12     # one evaluation / one step
TOL_RUN = 1e-9     # after N steps


def _eos(hs, kind):
    if kind == "default":
        return (hs.Barton2009(), hs.Barton2009())
    return (hs.Barton2009(), hs.Barton2009(_rho0=8.93, _c0=6.22, _cv=9.0e-4, _t0=300, _b0=3.16, _alpha=1, _beta=3.577, _gamma=2.088))


def _oeos(oracle, kind):
    if kind == "default":
        return [oracle.barton2009(), oracle.barton2009()]
    return [oracle.barton2009(), oracle.barton2009(c0=6.22, cv=9.0e-4, b0=3.16, beta=3.577, gamma=2.088)]


@pytest.mark.parametrize("kind", ["default", "hetero"])
def test_cell_functions_mph(gpu, oracle, kind):
    hs = gpu
    rng = np.random.default_rng(11)
    P = random_mph_prims(rng, 257)
    eos, oe = _eos(hs, kind), _oeos(oracle, kind)
    Q = hs.prim2cons_mph(eos, P)
    Qo, _ = oracle.prim2cons(oe, oracle.MPH30, P)
    assert relerr(Q, Qo) < TOL_OP
    P2 = hs.cons2prim_mph(eos, Qo)
    P2o, _ = oracle.cons2prim(oe, oracle.MPH30, Qo)
    assert relerr(P2, P2o) < 1e-11          # entropy = cv*log(S') amplifies roundoff of e_int by 1/(cv t0)
    F = hs.flux_mph(eos, Qo)
    Fo, _ = oracle.flux(oe, oracle.MPH30, Qo)
    assert relerr(F, Fo) < TOL_OP
    col = hs.noncons_flux(eos, Qo, dense=False)
    colo, _ = oracle.noncons_cols(oe, Qo)
    assert relerr(col, colo) < TOL_OP
    B = hs.noncons_flux(eos, Qo[:3])
    Bo = oracle.noncons_dense(oe, Qo[:3])
    assert B.shape == (3, 30, 30) and relerr(B.reshape(3, -1), Bo.reshape(3, -1), per_var=False) < TOL_OP
    eg = hs.get_eigvals(eos, Qo)
    ego, _ = oracle.get_eigvals(oe, oracle.MPH30, Qo)
    assert relerr(eg, ego, per_var=False) < TOL_OP


def test_cell_functions_sp(gpu, oracle):
    hs = gpu
    rng = np.random.default_rng(12)
    P = random_sp_prims(rng, 130)
    eos = hs.Barton2009()
    oe = [oracle.barton2009()]
    Q = hs.hyperelasticity.prim2cons(eos, P)
    Qo, _ = oracle.prim2cons(oe, oracle.SP13, P)
    assert relerr(Q, Qo) < TOL_OP
    F = hs.hyperelasticity.flux(eos, Qo)
    Fo, _ = oracle.flux(oe, oracle.SP13, Qo)
    assert relerr(F, Fo) < TOL_OP
    P2 = hs.hyperelasticity.cons2prim(eos, Qo)
    P2o, _ = oracle.cons2prim(oe, oracle.SP13, Qo)
    assert relerr(P2, P2o) < 1e-11
    eg = hs.hyperelasticity.get_eigvals(eos, Qo)
    ego, _ = oracle.get_eigvals(oe, oracle.SP13, Qo)
    assert relerr(eg, ego, per_var=False) < TOL_OP


@pytest.mark.parametrize("kind", ["default", "hetero"])
def test_hll_lxf_faces(gpu, oracle, kind):
    hs = gpu
    rng = np.random.default_rng(13)
    eos, oe = _eos(hs, kind), _oeos(oracle, kind)
    n = 65
    Pl = random_mph_prims(rng, n, spread=0.03); Pr = random_mph_prims(rng, n, spread=0.03)
    Ql, _ = oracle.prim2cons(oe, 1, Pl); Qr, _ = oracle.prim2cons(oe, 1, Pr)
    el, _ = oracle.get_eigvals(oe, 1, Ql); er, _ = oracle.get_eigvals(oe, 1, Qr)
    cons, dm, dp, s = hs.hll(eos, Ql, Qr, [el, er], return_speeds=True)
    co, dmo, dpo, so, st = oracle.hll(oe, Ql, Qr, el, er)
    assert st == 0
    assert np.all(cons == 0.0)                       # NumFluxes.jl:82
    assert relerr(s, so, per_var=False) < TOL_OP
    assert relerr(dm, dmo) < TOL_OP and relerr(dp, dpo) < TOL_OP
    lam = 11.0
    cons, dm, dp = hs.lxf(eos, Ql, Qr, lam)
    co, dmo, dpo, st = oracle.lxf(oe, Ql, Qr, lam)
    assert relerr(cons, co) < TOL_OP and relerr(dm, dmo) < TOL_OP and relerr(dp, dpo) < TOL_OP
    assert np.array_equal(dm, dp)                    # NumFluxes.jl:50-51


def test_golden_face_b4(gpu, oracle):
    """one HLL face of test case 6."""
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    Ql, Qr = hs.initial_states(eos, 6)
    eg = hs.get_eigvals(eos, np.stack([Ql, Qr]))
    _, dm, dp, s = hs.hll(eos, Ql, Qr, [eg[0], eg[1]], return_speeds=True)
    assert abs(s[0] - (-5.746359242338784)) < 1e-13 and abs(s[1] - 5.668441896578543) < 1e-13
    ref_dm = [-2.2136008888027487, -20.265569210477256, -5.7654714850916102, 1.1735829960630739, 1.8326561339193024, 3.9389545409005038]
    ref_dp16 = [-2.3511677908532080, -20.783870213302766, -9.5444778788371067, -11.105551423375545, -22.660677109939218, -48.102868115282782]
    assert np.allclose(dm[:6], ref_dm, rtol=1e-12, atol=0)
    assert np.allclose(dp[15:21], ref_dp16, rtol=1e-12, atol=0)


@pytest.mark.parametrize("flux", ["hll", "lxf"])
@pytest.mark.parametrize("tc,kind", [(6, "default"), (5, "default"), (7, "hetero")])
def test_run_mph_small(gpu, oracle, flux, tc, kind):
    hs = gpu
    eos, oe = _eos(hs, kind), _oeos(oracle, kind)
    nx, nsteps = 200, 25
    Ql, Qr = hs.initial_states(eos, tc)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    fk = oracle.HLL if flux == "hll" else oracle.LXF
    ref = oracle.run(oe, oracle.MPH30, fk, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=8)
    assert ref["status"] == 0
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0)
        # one step
        dt1 = sol.step(flux, 0.6, 1.0 / nx)
        one = oracle.run(oe, oracle.MPH30, fk, Q0, 0.6, 1.0 / nx, 1e9, 1, nthreads=8)
        assert abs(dt1[0] - one["dt"][0, 0]) <= 1e-13 * dt1[0]
        assert relerr(sol.download(), one["Q"]) < TOL_OP
        # N steps through the device-resident loop
        sol.upload(Q0)
        hist = sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps, record_dt=True)
        Q = sol.download()
        assert sol.steps[0] == nsteps
        assert np.allclose(hist[0], ref["dt"][0], rtol=1e-11, atol=0)
        assert abs(sol.t[0] - ref["t"][0]) < 1e-11 * ref["t"][0]
        assert relerr(Q, ref["Q"]) < TOL_RUN
        # frozen boundary cells (main.jl:219-220)
        assert np.array_equal(Q[0], Q0[0]) and np.array_equal(Q[-1], Q0[-1])


def test_golden_run_b5(gpu):
    """tc6, nx=16, 5 steps."""
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    Ql, Qr = hs.initial_states(eos, 6)
    with hs.Solver(eos, 16) as sol:
        sol.upload(hs.initial_condition(Ql, Qr, 16))
        hist = sol.advance(1e9, "hll", 0.6, 1 / 16, max_steps=5, record_dt=True)
        Q = sol.download()
    ref_dt = [0.006525871150502139, 0.006525871150502139, 0.006446700759844828, 0.005958085061464478, 0.005706267106528782]
    assert np.allclose(hist[0], ref_dt, rtol=1e-12, atol=0)
    assert np.allclose(Q[7, :6], [0.38297853292399647, 3.5431983793351409, 1.1321498895965438, 0.32995939011061942, 0.78173554155610159, 1.7603676372817543], rtol=1e-11)
    assert np.allclose(Q[:, [0, 1, 15, 16, 20]].sum(0), [7.9024417665803668, 71.345306122448974, 8.0975582334196332, 72.507755102040832, 136.32929234855314], rtol=1e-12)


@pytest.mark.parametrize("flux", ["hll", "lxf"])
def test_run_sp_small(gpu, oracle, flux):
    hs = gpu
    eos = hs.Barton2009()
    oe = [oracle.barton2009()]
    nx, nsteps = 300, 30
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    fk = oracle.HLL if flux == "hll" else oracle.LXF
    ref = oracle.run(oe, oracle.SP13, fk, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=8)
    with hs.Solver(eos, nx, model=hs.SP13) as sol:
        sol.upload(Q0)
        hist = sol.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps, record_dt=True)
        Q = sol.download()
    assert np.allclose(hist[0], ref["dt"][0], rtol=1e-11, atol=0)
    assert relerr(Q, ref["Q"]) < TOL_RUN


def test_ensemble_per_problem_dt(gpu, oracle):
    """Independent problems with their own dt / t (BASELINE config 4 in miniature), including
    problems that reach t_end at different step counts."""
    hs = gpu
    rng = np.random.default_rng(5)
    eos = (hs.Barton2009(), hs.Barton2009()); oe = [oracle.barton2009()] * 2
    nprob, nx = 6, 70
    Pl = random_mph_prims(rng, nprob, spread=0.03); Pr = random_mph_prims(rng, nprob, spread=0.03)
    Ql = hs.prim2cons_mph(eos, Pl); Qr = hs.prim2cons_mph(eos, Pr)
    Q0 = np.stack([hs.initial_condition(Ql[i], Qr[i], nx) for i in range(nprob)])
    t_end = 0.012
    ref = oracle.run(oe, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, t_end, 400, nthreads=8)
    assert ref["status"] == 0 and len(set(ref["steps"].tolist())) > 1
    with hs.Solver(eos, nx, nprob=nprob) as sol:
        sol.upload(Q0)
        sol.advance(t_end, "hll", 0.6, 1.0 / nx, max_steps=400)
        Q = sol.download()
        assert np.array_equal(sol.steps, ref["steps"])
        assert np.allclose(sol.t, ref["t"], rtol=1e-11)
    assert relerr(Q.reshape(-1, 30), ref["Q"].reshape(-1, 30)) < TOL_RUN


def test_domain_error(gpu):
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    Ql, _ = hs.initial_states(eos, 6)
    bad = Ql.copy(); bad[6:15] *= -1.0       # det(F) < 0 -> sqrt(negative) in cons2prim (HyperelasticityMPh.jl:114)
    with pytest.raises(hs.DomainError):
        hs.cons2prim_mph(eos, bad)


def test_driver_default_run_matches_oracle(gpu, oracle, tmp_path):
    """The shipped default configuration in miniature through the main.jl-compatible driver:
    CSV snapshots every log_freq steps, result.csv, restart from the last snapshot."""
    import logging
    from hyperelasticsolver_b200 import driver
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    d = str(tmp_path / "barton_data") + "/"
    nx, T = 120, 0.012
    Q, t, steps = driver.run(eos, 6, nx, 0.6, T, 1.0, 10, d, "hll", 0, None, logging.getLogger("test"))
    Ql, Qr = hs.initial_states(eos, 6)
    ref = oracle.run(None, oracle.MPH30, oracle.HLL, hs.initial_condition(Ql, Qr, nx), 0.6, 1.0 / nx, T, 10000, nthreads=8)
    assert steps == ref["steps"][0] and abs(t - ref["t"][0]) < 1e-12 * T
    assert relerr(Q, ref["Q"]) < 1e-9
    import os
    files = sorted(os.listdir(d))
    assert files[0] == "result.csv" and "sol_000000.csv" in files and f"sol_{(steps // 10) * 10:06d}.csv" in files
    P, n = driver.read_data(os.path.join(d, "result.csv"))
    Po, _ = oracle.cons2prim(None, oracle.MPH30, ref["Q"])
    assert n == nx and relerr(P, Po) < 1e-8
    # tanh-smoothed initial condition (main.jl:110-123): alpha profile and alpha1 + alpha2 = 1
    Q0 = driver.initial_condition_tanh(eos, Ql, Qr, nx, 0.2)
    x = (np.arange(1, nx + 1) - 0.5) / nx
    assert np.allclose(Q0[:, 0], 0.1 * (np.tanh(4 * (x - 0.5) / 0.2) + 1) + 0.4, rtol=1e-15)
    assert np.allclose(Q0[:, 0] + Q0[:, 15], 1.0, rtol=0, atol=1e-16)


def test_config0_default_run_641_steps(gpu, oracle):
    """BASELINE config 0 = main.jl as shipped: two-phase, test case 6, nx = 1000, cfl 0.6, T = 0.06,
    HLL.  SURVEY.md B.6: 641 steps, t_end = 0.06003958139325407; parity <= 1e-9 after N steps."""
    import json, os
    hs = gpu
    B = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_appendix_b.json")))["default_run"]
    eos = (hs.Barton2009(), hs.Barton2009())
    nx, T = 1000, 0.06
    Ql, Qr = hs.initial_states(eos, 6)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0)
        hist = sol.advance(T, "hll", 0.6, 1.0 / nx, max_steps=1000, record_dt=True)
        Q = sol.download()
        steps, t = int(sol.steps[0]), float(sol.t[0])
    assert steps == B["steps"] == 641
    assert abs(t - B["t_end"]) < 1e-12 * t
    dts = hist[0, :steps]
    assert abs(dts.min() - B["dt_min"]) < 1e-11 * B["dt_min"] and abs(dts.max() - B["dt_max"]) < 1e-11 * B["dt_max"]
    assert abs(dts[-1] - B["dt_last"]) < 1e-10 * B["dt_last"]
    assert np.allclose(Q[500, :6], B["cell501_Q_1_6"], rtol=1e-10)
    assert Q[:, 0].min() >= 0.1 - 1e-12 and Q[:, 0].max() <= 0.9 + 1e-12 and np.abs(Q[:, 0] + Q[:, 15] - 1).max() < 1e-14
    ref = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, T, 1000, nthreads=oracle.hardware_threads())
    assert ref["steps"][0] == 641 and relerr(Q, ref["Q"]) < 1e-9
    assert np.allclose(dts, ref["dt"][0, :641], rtol=1e-11, atol=0)


@pytest.mark.parametrize("model", ["mph30", "sp13"])
def test_config4_ensemble_sample_of_64(gpu, oracle, model):
    """BASELINE config 4 in miniature with its own generator (bench.ensemble_states): 64 random
    problems validated end to end against the oracle, per-problem dt."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from hyperelasticsolver_b200.slab import CudaKernels, EnsembleSolver
    hs = gpu
    nprob, nx, nsteps = 64, 192, 16
    eos, hmodel, Qlr = bench.ensemble_states(hs, model, 1000, 1000 + nprob)     # problems 1000..1063 of the 65,536
    om = oracle.MPH30 if model == "mph30" else oracle.SP13
    oe = [oracle.barton2009()] * (2 if model == "mph30" else 1)
    Q0 = np.stack([hs.initial_condition(Qlr[i, 0], Qlr[i, 1], nx) for i in range(nprob)])
    sol = EnsembleSolver(CudaKernels(eos, hmodel, "cuda:0"), nx, nprob)
    sol.set_local(Q0)
    for _ in range(nsteps):
        sol.step(hs.HLL, 0.6, 1.0 / nx)
    sol.check_status()
    ref = oracle.run(oe, om, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=oracle.hardware_threads())
    assert ref["status"] == 0                                  # admissibility of the generator's states
    nv = Q0.shape[-1]
    assert relerr(sol.local().reshape(-1, nv), ref["Q"].reshape(-1, nv)) < 1e-9
    assert np.allclose(sol.t, ref["t"], rtol=1e-11) and len(set(np.round(sol.t, 12))) > 32    # genuinely different dt per problem


def test_get_eigvals_general_normal_and_full_sweep(gpu, oracle):
    """the physics takes a normal n (EquationsOfState.jl:223, HyperelasticityMPh.jl:264);
    plus the full-spectrum CFL sweep of a resident grid (hs_wave_speeds with eig != NULL)."""
    hs = gpu
    rng = np.random.default_rng(21)
    eos = (hs.Barton2009(), hs.Barton2009()); oe = [oracle.barton2009()] * 2
    Q = hs.prim2cons_mph(eos, random_mph_prims(rng, 40))
    for n in ([0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [0.6, 0.0, 0.8], list(np.array([1.0, -2.0, 0.5]) / np.linalg.norm([1.0, -2.0, 0.5]))):
        eg = hs.get_eigvals(eos, Q, n)
        ego, st = oracle.get_eigvals_n(oe, Q, n)
        assert st == 0 and relerr(eg, ego, per_var=False) < 1e-12
    with pytest.raises(hs.HyperelasticError):
        hs.get_eigvals(eos, Q, [2.0, 0.0, 0.0])          # not a unit vector
    # anchor: F = I, S = 0 -> speeds (b0, b0, c0) for any direction (isotropy of the reference state)
    P = np.zeros(30)
    for p in range(2):
        P[15 * p:15 * p + 15] = [0.5, 8.93, 0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1]
    e0 = hs.get_eigvals(eos, hs.prim2cons_mph(eos, P), [0.0, 0.6, 0.8])
    assert np.allclose(e0[:3], [2.1, 2.1, 4.6], rtol=1e-12) and np.allclose(e0[3:6], [-2.1, -2.1, -4.6], rtol=1e-12)
    # full sweep on a resident grid
    nx = 77
    Ql, Qr = hs.initial_states(eos, 7)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0); sol.step()
        lam, eig = sol.wave_speeds(full=True)
        Q1 = sol.download()
    ego, _ = oracle.get_eigvals(oe, oracle.MPH30, Q1)
    assert relerr(eig, ego, per_var=False) < 1e-12 and abs(lam[0] - np.abs(ego).max()) < 1e-12 * lam[0]


def test_device_math_selftest(gpu):
    """The hot path's branch-free 1/x, 1/sqrt(x), sqrt(x) (hardware seed + two Newton FMAs) and its Newton
    largest-eigenvalue solve, evaluated ON THE DEVICE, against numpy."""
    from hyperelasticsolver_b200 import _lib as L
    rng = np.random.default_rng(8)
    x = np.concatenate([10.0 ** rng.uniform(-6, 6, 200000), rng.uniform(0.01, 100.0, 200000), [1.0, 2.0, 4.0, 0.25, 1e-300, 1e300]])
    rcp = np.empty_like(x); rsq = np.empty_like(x); sq = np.empty_like(x)
    L.check(L.lib().hs_selftest_math(x.ctypes.data, rcp.ctypes.data, rsq.ctypes.data, sq.ctypes.data, x.size, 0))
    ulp = lambda got, ref: np.abs(got - ref) / np.spacing(np.abs(ref))
    assert ulp(rcp, 1.0 / x).max() <= 2.0
    assert ulp(rsq, 1.0 / np.sqrt(x)).max() <= 2.0
    assert ulp(sq, np.sqrt(x)).max() <= 2.0
    z = np.zeros(1); o = [np.empty(1) for _ in range(3)]
    L.check(L.lib().hs_selftest_math(z.ctypes.data, o[0].ctypes.data, o[1].ctypes.data, o[2].ctypes.data, 1, 0))
    assert o[2][0] == 0.0                                     # sqrt(0) = 0
    # largest |eigenvalue|: positive definite, degenerate pairs, indefinite
    n = 100000
    S6 = np.empty((n, 6)); ref = np.empty(n)
    Qs = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    ev = np.sort(rng.uniform(0.5, 30.0, (n, 3)), axis=1)
    ev[::3, 1] = ev[::3, 0] * (1 + rng.uniform(0, 1e-6, ev[::3, 0].shape))                 # degenerate shear pair
    ev[::7, 1] = ev[::7, 2] * (1 - 10.0 ** rng.uniform(-12, -1, ev[::7, 2].shape))          # (near-)degenerate largest pair
    ev[::11] -= rng.uniform(0, 40.0, (ev[::11].shape[0], 1))                              # indefinite
    M = np.einsum("nij,nj,nkj->nik", Qs, ev, Qs); M = 0.5 * (M + M.transpose(0, 2, 1))
    S6[:] = np.stack([M[:, 0, 0], M[:, 0, 1], M[:, 0, 2], M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]], axis=1)
    ref = np.abs(np.linalg.eigvalsh(M)).max(axis=1)
    out = np.empty(n)
    L.check(L.lib().hs_selftest_eig(S6.ctypes.data, out.ctypes.data, n, 0))
    assert (np.abs(out - ref) / ref).max() < 5e-15


@pytest.mark.parametrize("tc", [1, 2, 3, 4, 5, 6, 7, 10])
def test_all_shipped_test_cases(gpu, oracle, tc):
    """Every Riemann problem of initial_states (HyperelasticityMPh.jl:276-408): first dt against
    SURVEY.md B.3, then 15 HLL steps against the oracle (tc 1-2 are strongly pre-strained: density 5.0
    against rho0 = 8.93; tc 1-5 have identical phases, which must stay bit-identical with alpha = 1/2)."""
    import json, os
    hs = gpu
    B = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "survey_appendix_b.json")))["lambda_max_dt_nx1000_cfl06"]
    eos = (hs.Barton2009(), hs.Barton2009())
    nx = 1000
    Ql, Qr = hs.initial_states(eos, tc)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx) as sol:
        sol.upload(Q0)
        lam = sol.wave_speeds()[0]
        assert abs(lam - B[str(tc)][0]) < 1e-13 * lam
        dt = sol.step("hll", 0.6, 1.0 / nx)[0]
        assert abs(dt - B[str(tc)][1]) < 1e-13 * dt
        sol.upload(Q0)
        sol.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=15)
        Q = sol.download()
    ref = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, 15, nthreads=oracle.hardware_threads())
    assert ref["status"] == 0 and relerr(Q, ref["Q"]) < 1e-9
    if tc <= 5:
        assert np.array_equal(Q[:, :15], Q[:, 15:]) and np.all(Q[:, 0] == 0.5)


def test_c_abi_from_plain_c(gpu, tmp_path):
    """examples/abi_smoke.c: the library driven from C (no Python / torch in the process), golden run B.5."""
    import os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.join(root, "hyperelasticsolver_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(root, "include"), os.path.join(root, "examples", "abi_smoke.c"), "-o", exe,
                           "-L" + libdir, "-lhyperelastic_b200", "-Wl,-rpath," + libdir, "-lm"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ABI-SMOKE-OK" in out.stdout, out.stdout + out.stderr
