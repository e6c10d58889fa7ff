"""CPU: pin the C++ oracle against (i) SURVEY.md Appendix B, (ii) the committed vectors of the
independent torch-autograd restatement, (iii) physical anchors.  (The reference ships no golden
vectors and Julia is unavailable -> "parity unpinned", see oracle/README.md.)"""
import json
import os

import numpy as np
import pytest

from hyperelasticsolver_b200.testcases import mph_primitive_states, riemann_grid, sp_primitive_states

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def B():
    return json.load(open(os.path.join(G, "survey_appendix_b.json")))


def _states(oracle, tc):
    Pl, Pr = mph_primitive_states(tc)
    Q, st = oracle.prim2cons(None, oracle.MPH30, np.stack([Pl, Pr]))
    assert st == 0
    return Q[0], Q[1]


def test_quadrature(oracle, B):
    x, w = oracle.quadrature(False)
    assert np.allclose(x, B["gl6_nodes"], rtol=0, atol=3e-16) and np.allclose(w, B["gl6_weights"], rtol=0, atol=3e-16)
    x, w = oracle.quadrature(True)
    assert np.allclose(x, B["lobatto6_nodes"], rtol=0, atol=5e-16) and np.allclose(w, B["lobatto6_weights"], rtol=0, atol=5e-16)  # survey values carry the ulp noise of a computed rule
    assert abs(w.sum() - 1) < 1e-15


def test_tc6_states_and_speeds(oracle, B):
    Ql, Qr = _states(oracle, 6)
    assert np.allclose(Ql, B["tc6_Ql"], rtol=1e-15, atol=0)
    assert np.allclose(Qr[:15], B["tc6_Qr_phase1"], rtol=4e-16, atol=0)
    assert np.allclose(Qr[15:], B["tc6_Qr_phase2"], rtol=1e-15, atol=0)
    P, _ = oracle.cons2prim(None, oracle.MPH30, Ql)
    assert np.allclose(P[:15], B["tc6_cons2prim_Ql_phase1"], rtol=2e-15, atol=0)
    eg, _ = oracle.get_eigvals(None, oracle.MPH30, np.stack([Ql, Qr]))
    u1 = P[2]
    assert np.allclose(np.sort(eg[0, :3] - u1)[::-1], B["tc6_c_left"], rtol=1e-14)
    assert np.allclose(np.sort(eg[1, :3])[::-1], B["tc6_c_right"], rtol=1e-14)


def test_lambda_max_all_cases(oracle, B):
    for tc, (lam, dt) in B["lambda_max_dt_nx1000_cfl06"].items():
        Ql, Qr = _states(oracle, int(tc))
        l = oracle.lambda_max(None, oracle.MPH30, np.stack([Ql, Qr]))
        assert abs(l - lam) < 1e-14 * lam
        assert abs(0.6 * 1e-3 / l - dt) < 1e-14 * dt


def test_hll_face(oracle, B):
    Ql, Qr = _states(oracle, 6)
    eg, _ = oracle.get_eigvals(None, oracle.MPH30, np.stack([Ql, Qr]))
    cons, dm, dp, s, st = oracle.hll(None, Ql, Qr, eg[0], eg[1])
    h = B["hll_face_tc6"]
    assert st == 0 and np.all(cons == 0)
    assert abs(s[0, 0] - h["s_l"]) < 1e-14 and abs(s[0, 1] - h["s_r"]) < 1e-13
    for got, key in ((dm[0, :6], "dm_1_6"), (dm[0, 15:21], "dm_16_21"), (dp[0, :6], "dp_1_6"), (dp[0, 15:21], "dp_16_21")):
        assert np.allclose(got, h[key], rtol=2e-14, atol=0)


def test_small_run(oracle, B):
    Ql, Qr = _states(oracle, 6)
    r = oracle.run(None, oracle.MPH30, oracle.HLL, riemann_grid(Ql, Qr, 16), 0.6, 1 / 16, 1e9, 5)
    g = B["run_tc6_nx16_5steps"]
    assert np.allclose(r["dt"][0], g["dt"], rtol=1e-14, atol=0)
    assert np.allclose(r["Q"][7, :6], g["cell8_Q_1_6"], rtol=1e-14) and np.allclose(r["Q"][8, :6], g["cell9_Q_1_6"], rtol=1e-14)
    assert np.allclose(r["Q"][:, [0, 1, 15, 16, 20]].sum(0), g["colsum_Q1_Q2_Q16_Q17_Q21"], rtol=1e-14)
    # literal mode (update_cell per cell: every face twice, main.jl:43-60) is bit-identical
    r2 = oracle.run(None, oracle.MPH30, oracle.HLL, riemann_grid(Ql, Qr, 16), 0.6, 1 / 16, 1e9, 5, literal=True)
    assert np.array_equal(r["Q"], r2["Q"])


def test_against_pyoracle_vectors(oracle):
    d = json.load(open(os.path.join(G, "pyoracle_vectors.json")))
    rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()
    assert len(d["cases"]) >= 12
    for c in d["cases"]:
        eos = [np.array(b) for b in c["eos_blocks"]]
        Q, _ = oracle.prim2cons(eos, 1, np.array(c["P"]))
        assert rel(Q, c["Q"]) < 1e-14
        Q = np.array(c["Q"])
        assert rel(oracle.cons2prim(eos, 1, Q)[0], c["cons2prim"]) < 1e-13
        assert rel(oracle.flux(eos, 1, Q)[0], c["flux"]) < 1e-13
        assert rel(oracle.get_eigvals(eos, 1, Q)[0][0], c["eigvals"]) < 1e-13
        assert rel(oracle.noncons_cols(eos, Q)[0], c["noncons_cols"]) < 1e-13
        for p in range(2):  # asymmetry of the reference's acoustic tensor is roundoff only
            ac = np.array(c["acoustic"][p])
            assert np.abs(ac - ac.T).max() < 1e-14 * np.abs(ac).max()


def test_physical_anchors(oracle):
    """F = I, S = 0 -> zero stress, T = t0, speeds (b0, b0, c0)."""
    for e in (oracle.barton2009(), oracle.barton2009(c0=6.22, cv=9.0e-4, b0=3.16, beta=3.577, gamma=2.088)):
        I9 = np.eye(3).flatten()
        assert np.abs(oracle.stress(e, 0.0, I9)).max() < 1e-12
        assert abs(oracle.temperature(e, 0.0, I9) - e[3]) < 1e-10
        ev = np.linalg.eigvalsh(oracle.acoustic(e, 0.0, I9))
        assert np.allclose(np.sqrt(ev), [e[4], e[4], e[1]], rtol=1e-13)


def test_prim_cons_roundtrip_and_identical_phases(oracle):
    """prim2cons o cons2prim == id on Q; identical phases stay bit-identical (tc 1-5)."""
    Ql, Qr = _states(oracle, 5)
    P, _ = oracle.cons2prim(None, 1, Ql)
    Q2, _ = oracle.prim2cons(None, 1, P)
    m = np.ones(30, bool); m[[1, 16]] = False     # Q[2] is passive (quirk Q2): rho comes from det
    assert np.allclose(Q2[m], Ql[m], rtol=1e-13)
    r = oracle.run(None, 1, oracle.HLL, riemann_grid(Ql, Qr, 40), 0.6, 1 / 40, 1e9, 12)
    assert np.array_equal(r["Q"][:, :15], r["Q"][:, 15:])
    assert np.all(r["Q"][:, 0] == 0.5)


def test_conservation_tc6(oracle):
    """per-phase mass and mixture momentum/energy change only by the boundary-flux budget."""
    Ql, Qr = _states(oracle, 6)
    nx, n = 64, 10
    Q0 = riemann_grid(Ql, Qr, nx)
    r = oracle.run(None, 1, oracle.HLL, Q0, 0.6, 1 / nx, 1e9, n)
    Q = r["Q"]; t = r["t"][0]
    Fl, _ = oracle.flux(None, 1, Ql); Fr, _ = oracle.flux(None, 1, Qr)
    interior = slice(1, nx - 1)   # boundary cells are frozen, interior obeys the conservation law
    for cols in ([1], [16], [2, 17], [3, 18], [4, 19], [5, 20]):
        tot0 = Q0[interior][:, cols].sum(); tot1 = Q[interior][:, cols].sum()
        budget = -(t * nx) * (Fr[cols].sum() - Fl[cols].sum())
        assert abs((tot1 - tot0) - budget) < 1e-10 * max(1.0, abs(tot0))


def sp_to_mph(Qsp):
    """two identical phases at alpha = 1/2 carrying Q_SP/2 each (F row-major -> column-major)."""
    Qsp = np.asarray(Qsp)
    n = Qsp.shape[0]
    ph = np.zeros((n, 15))
    ph[:, 0] = 0.5
    ph[:, 2:5] = 0.5 * Qsp[:, 0:3]
    ph[:, 5] = 0.5 * Qsp[:, 12]
    ph[:, 6:15] = 0.5 * Qsp[:, 3:12].reshape(n, 3, 3).transpose(0, 2, 1).reshape(n, 9)
    return np.concatenate([ph, ph], axis=1)


def test_sp_matches_mph_identical_phases(oracle):
    """the shipped two-phase algorithm with two identical phases at alpha = 1/2
    reduces to the one-phase model: Q_mph[phase] = Q_SP / 2 (F column- vs row-major)."""
    Pl, Pr = sp_primitive_states(1)
    Qs, _ = oracle.prim2cons(None, oracle.SP13, np.stack([Pl, Pr]))
    nx, n = 60, 20
    Q0 = riemann_grid(Qs[0], Qs[1], nx)
    rs = oracle.run(None, oracle.SP13, oracle.HLL, Q0, 0.6, 1 / nx, 1e9, n)
    rm = oracle.run(None, oracle.MPH30, oracle.HLL, sp_to_mph(Q0), 0.6, 1 / nx, 1e9, n)
    assert np.allclose(rs["dt"][0], rm["dt"][0], rtol=1e-13)
    got = rm["Q"].copy(); want = sp_to_mph(rs["Q"])
    got[:, [1, 16]] = 0.0     # alpha*rho is passive and absent from the 13-variable model
    assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()


def test_hank2016_oracle_anchors(oracle):
    """Hank2016 (EquationsOfState.jl:301-356) is dead code without tests in the reference: the restatement is
    anchored on what the formulas imply -- zero elastic energy and stress in the undeformed state, a trace-free
    stress, the small-strain shear response sigma = -2 (den/rho0) mu eps for distortion 1 + eps, pressure as the
    inverse of energy, and the dual-number gradient against central differences."""
    eos = oracle.hank2016()
    rho0, mu, gamma, pinf, a = eos
    I9 = np.eye(3).flatten()
    e, st = oracle.hank_energy(eos, rho0, 1e9, I9)
    assert st == 0 and abs(e - (1e9 + gamma * pinf) / (rho0 * (gamma - 1))) < 1e-6
    s, _ = oracle.hank_stress(eos, rho0, 1e9, I9)
    assert np.abs(s).max() < 1e-14 * mu                                    # roundoff of mu
    rng = np.random.default_rng(11)
    for it in range(20):
        A = np.eye(3) + 0.15 * rng.uniform(-1, 1, (3, 3)); a9 = A.flatten(order="F")
        den = rho0 * np.linalg.det(A)
        s, st = oracle.hank_stress(eos, den, 3e9, a9); S = s.reshape(3, 3, order="F")
        assert st == 0 and abs(np.trace(S)) < 1e-12 * np.abs(S).max()
        assert np.abs(S - S.T).max() < 1e-12 * np.abs(S).max()
        # central differences of energy over the 9 entries of G, then -2 den G de/dG
        G = (A.T @ A)
        g9 = G.flatten(order="F"); h = 1e-6; d = np.zeros(9)
        for k in range(9):
            gp = g9.copy(); gm = g9.copy(); gp[k] += h; gm[k] -= h
            d[k] = (oracle.hank_energy(eos, den, 3e9, gp)[0] - oracle.hank_energy(eos, den, 3e9, gm)[0]) / (2 * h)
        fd = -2 * den * G @ d.reshape(3, 3, order="F")
        assert np.abs(fd - S).max() < 2e-5 * np.abs(S).max()
        g_inv = oracle.invariants(g9)
        e, _ = oracle.hank_energy(eos, den, 3e9, g9)
        assert abs(oracle.hank_pressure(eos, den, e, g_inv)[0] - 3e9) < 1e-12 * gamma * pinf
    eps = 1e-7 * np.array([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 0.0]])
    s, _ = oracle.hank_stress(eos, rho0, 0.0, (np.eye(3) + eps).flatten(order="F"))
    assert abs(s.reshape(3, 3)[0, 1] - (-2 * mu * 1e-7)) < 1e-6 * 2 * mu * 1e-7


def test_hank2016_against_pyoracle_vectors(oracle):
    """C++ dual-number Hank2016 vs the torch reverse-mode restatement (tests/golden/pyoracle_hank_vectors.json)"""
    d = json.load(open(os.path.join(G, "pyoracle_hank_vectors.json")))
    assert len(d["cases"]) >= 12
    for c in d["cases"]:
        eos = np.array(c["eos_block"])
        e, st = oracle.hank_energy(eos, c["den"], c["pres"], np.array(c["G"]))
        assert st == 0 and abs(e - c["energy"]) <= 1e-13 * abs(c["energy"])
        assert np.abs(oracle.invariants(np.array(c["G"])) - np.array(c["invariants"])).max() <= 1e-13 * max(np.abs(c["invariants"]))
        p, st = oracle.hank_pressure(eos, c["den"], c["energy"], np.array(c["invariants"]))
        assert st == 0 and abs(p - c["pressure"]) <= 1e-12 * eos[2] * eos[3]
        s_, st = oracle.hank_stress(eos, c["den"], c["pres"], np.array(c["distortion"]))
        assert st == 0 and np.abs(s_ - np.array(c["stress"])).max() <= 1e-12 * np.abs(c["stress"]).max()


def _random_mph_states(oracle, rng, n, eos=None):
    P = []
    for _ in range(n):
        a1 = rng.uniform(0.1, 0.9); row = []
        for a in (a1, 1 - a1):
            F = np.eye(3) + 0.1 * rng.uniform(-1, 1, (3, 3))
            row += [a, 8.9 / np.linalg.det(F), *rng.uniform(-1, 1, 3), rng.uniform(0, 1e-3), *F.flatten(order="F")]
        P.append(row)
    Q, st = oracle.prim2cons(eos, 1, np.array(P))
    assert st == 0
    return Q


def test_numerical_flux_consistency(oracle):
    """Properties every consistent path-conservative flux has, independent of any golden number (NumFluxes.jl:25-132):
    equal states give no fluctuation and (LxF) the physical flux; D- + D+ of LxF is the whole path integral; swapping the
    two phases of both states swaps the two halves of every output."""
    rng = np.random.default_rng(21)
    Q = _random_mph_states(oracle, rng, 6)
    eig, _ = oracle.get_eigvals(None, 1, Q)
    F, _ = oracle.flux(None, 1, Q)
    cons, dm, dp, s, st = oracle.hll(None, Q, Q, eig, eig)
    assert st == 0 and np.all(cons == 0.0)
    assert np.abs(dm).max() < 1e-12 * np.abs(F).max() and np.abs(dp).max() < 1e-12 * np.abs(F).max()
    lam = 7.0
    cons, dm, dp, st = oracle.lxf(None, Q, Q, lam)
    assert np.abs(cons - F).max() < 1e-13 * np.abs(F).max() and np.all(dm == 0.0) and np.all(dp == 0.0)
    Ql, Qr = Q[:3], Q[3:]
    cons, dm, dp, st = oracle.lxf(None, Ql, Qr, lam)
    Fl, Fr = F[:3], F[3:]
    assert np.abs(cons - (0.5 * (Fl + Fr) - 0.5 * lam * (Qr - Ql))).max() < 1e-13 * np.abs(F).max()
    assert np.array_equal(dm, dp)                                                    # NumFluxes.jl:50-51
    # phase-swap symmetry (k = (1/2, 1/2), omega = 0: nothing distinguishes the phases, HyperelasticityMPh.jl:205-217)
    sw = lambda X: np.concatenate([X[..., 15:], X[..., :15]], axis=-1)
    swe = lambda E: np.concatenate([E[..., 6:], E[..., :6]], axis=-1)
    c1, m1, p1, s1, _ = oracle.hll(None, Ql, Qr, eig[:3], eig[3:])
    c2, m2, p2, s2, _ = oracle.hll(None, sw(Ql), sw(Qr), swe(eig[:3]), swe(eig[3:]))
    scale = np.abs(m1).max()
    assert np.abs(sw(m2) - m1).max() < 1e-13 * scale and np.abs(sw(p2) - p1).max() < 1e-13 * scale and np.array_equal(s1, s2)


def test_wave_speeds_rotate_with_the_frame(oracle):
    """get_eigvals for a unit normal n (HyperelasticityMPh.jl:252-266, EquationsOfState.jl:223) equals get_eigvals for
    (1,0,0) of the state seen from a frame rotated by R with R n = e1: u -> R u, F -> R F (objectivity)."""
    rng = np.random.default_rng(23)
    Q = _random_mph_states(oracle, rng, 8)
    for k in range(8):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        a = np.cross(n, [1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.cross(n, [0.0, 1.0, 0.0])
        a /= np.linalg.norm(a)
        R = np.stack([n, a, np.cross(n, a)])                                          # rows: R n = e1, orthonormal
        assert np.allclose(R @ n, [1, 0, 0]) and np.allclose(R @ R.T, np.eye(3))
        q = Q[k].copy(); qr = q.copy()
        for p in (0, 15):
            qr[p + 2:p + 5] = R @ q[p + 2:p + 5]
            qr[p + 6:p + 15] = (R @ q[p + 6:p + 15].reshape(3, 3, order="F")).flatten(order="F")
        en, st = oracle.get_eigvals_n(None, q[None], n)
        e1, st1 = oracle.get_eigvals(None, 1, qr[None])
        assert st == 0 and st1 == 0
        assert np.abs(en - e1).max() < 1e-12 * np.abs(e1).max()


def test_oracle_within_roundoff_of_exact_arithmetic(oracle):
    """oracle/mporacle.py evaluates the reference's formulas in 50-digit arithmetic (nested dual numbers over mpmath):
    tests/golden/mp_vectors.json holds cons2prim / flux / acoustic tensor / eigvals / non-conservative columns of the 24
    golden phase states and two complete path-conservative HLL faces to 30 digits.  The FP64 C++ oracle must sit within
    a few 1e-15 of them (1e-14 .. 1e-13 where the reference's own formula cancels: flux differences, S = cv log(1 +
    (e - ...)/(cv t0 ...))): this bounds the oracle's roundoff ~100x below the 1e-12 parity tolerance of the GPU tests."""
    d = json.load(open(os.path.join(G, "mp_vectors.json")))
    f = lambda L: np.array([float(x) for x in L])
    rel = lambda a, b: np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()
    assert len(d["cases"]) == 12
    for c in d["cases"]:
        eos = [np.array(b) for b in c["eos_blocks"]]
        Q = np.array(c["Q"])
        P = oracle.cons2prim(eos, 1, Q)[0]
        Pm = f(c["cons2prim"])
        assert rel(P, Pm) < 5e-15
        for k in (5, 20):
            assert abs(P[k] - Pm[k]) < 1e-13 * abs(Pm[k])          # entropy: roundoff of e_int amplified by 1/(cv t0)
        assert rel(oracle.flux(eos, 1, Q)[0], f(c["flux"])) < 1e-13
        assert rel(oracle.get_eigvals(eos, 1, Q)[0][0], f(c["eigvals"])) < 1e-14
        assert rel(oracle.noncons_cols(eos, Q)[0], f(c["noncons_cols"])) < 5e-14
        for p in range(2):
            ac = oracle.acoustic(eos[p], P[15 * p + 5], P[15 * p + 6:15 * p + 15])
            am = np.array([[float(x) for x in row] for row in c["acoustic"][p]])
            assert np.abs(ac - am).max() < 2e-14 * np.abs(am).max()
            assert np.abs(am - am.T).max() < 1e-28 * np.abs(am).max()      # symmetric in exact arithmetic (30 digits stored)
    for fc in d["hll_faces"]:
        eos = [np.array(b) for b in fc["eos_blocks"]]
        Ql, Qr = np.array(fc["Ql"]), np.array(fc["Qr"])
        el, _ = oracle.get_eigvals(eos, 1, Ql[None]); er, _ = oracle.get_eigvals(eos, 1, Qr[None])
        _, dm, dp, s, st = oracle.hll(eos, Ql[None], Qr[None], el, er)
        assert st == 0
        assert abs(s[0][0] - float(fc["s_l"])) < 5e-15 * abs(s[0][0]) and abs(s[0][1] - float(fc["s_r"])) < 5e-15 * abs(s[0][1])
        assert rel(dm[0], f(fc["dm"])) < 1e-14 and rel(dp[0], f(fc["dp"])) < 1e-14
    x, w = oracle.quadrature(False)
    assert np.abs(x - f(d["gauss_legendre6"]["x"])).max() < 1e-16 and np.abs(w - f(d["gauss_legendre6"]["w"])).max() < 1e-16


def test_mp_restatement_anchors():
    """the arbitrary-precision restatement itself, on facts that need no other implementation: zero stress, T = t0 and speeds
    (b0, b0, c0) at F = I, S = 0 to 40 digits; derivative identities; regenerating one stored case reproduces the file"""
    import importlib.util
    import mpmath as mp
    spec = importlib.util.spec_from_file_location("mporacle", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "mporacle.py"))
    M = importlib.util.module_from_spec(spec); spec.loader.exec_module(M)
    blk = [8.93, 4.6, 3.9e-4, 300.0, 2.1, 1.0, 3.0, 2.0, 2.1 ** 2, 4.6 ** 2 - (4 / 3) * 2.1 ** 2]
    eos = M.Barton2009(blk)
    I9 = [mp.mpf(x) for x in (1, 0, 0, 0, 1, 0, 0, 0, 1)]
    assert max(abs(x) for x in M.stress(eos, mp.mpf(0), I9)) < mp.mpf("1e-40")
    T = M.energy(eos, M.Dual(mp.mpf(0), mp.mpf(1)), M.finger(I9)).d
    assert abs(T - eos.t0) < mp.mpf("1e-40")
    ac = M.acoustic(eos, mp.mpf(0), I9)
    ev = M.eig_sym3(ac)
    # longitudinal speed^2 = k0 + 4/3 b0sq with the struct's (rounded) k0, b0sq; shear speed^2 = b0sq
    assert abs(ev[2] - (eos.k0 + mp.mpf(4) / 3 * eos.b0sq)) < mp.mpf("1e-40") and abs(ev[0] - eos.b0sq) < mp.mpf("1e-40") and abs(ev[1] - eos.b0sq) < mp.mpf("1e-40")
    # entropy inverts energy in S
    F = [mp.mpf(x) for x in (0.98, 0.02, 0.0, 0.0, 1.0, 0.0, 0.0, 0.1, 1.0)]
    Gf = M.finger(F)
    S = mp.mpf("7.5e-4")
    assert abs(M.entropy(eos, M.energy(eos, S, Gf), Gf) - S) < mp.mpf("1e-45")
    d = json.load(open(os.path.join(G, "mp_vectors.json")))
    c = d["cases"][7]
    eoss = [M.Barton2009(b) for b in c["eos_blocks"]]
    Q = [mp.mpf(float(x)) for x in c["Q"]]
    again = [M.s30(x) for x in M.noncons_cols(eoss, Q)]
    assert again == c["noncons_cols"]


def test_oracle_time_step_within_roundoff_of_exact_arithmetic(oracle):
    """two complete passes of the loop body of main.jl:204-227 (CFL sweep, dt, frozen boundary cells, update_cell with the
    path-conservative hll on every face) restated independently in 50-digit arithmetic (oracle/mporacle.py::time_step):
    the C++ oracle's run -- in both of its modes -- reproduces dt and the state to roundoff"""
    d = json.load(open(os.path.join(G, "mp_vectors.json")))["two_steps"]
    eos = [np.array(b) for b in d["eos_blocks"]]
    Q0 = np.array(d["Q0"]); nx = d["nx"]
    want = np.array([[float(x) for x in c] for c in d["Q"]])
    dtw = np.array([float(x) for x in d["dt"]])
    for literal in (False, True):
        r = oracle.run(eos, oracle.MPH30, oracle.HLL, Q0, d["cfl"], 1.0 / nx, 1e9, 2, literal=literal)
        assert r["status"] == 0
        assert np.abs(r["dt"][0] - dtw).max() < 1e-15 * dtw.max()
        scale = np.maximum(np.abs(want).max(axis=0), 1e-3 * np.abs(want).max())
        assert (np.abs(r["Q"] - want).max(axis=0) / scale).max() < 5e-14        # measured 1.1e-14
        assert np.array_equal(r["Q"][0], Q0[0]) and np.array_equal(r["Q"][-1], Q0[-1])      # main.jl:219-220
