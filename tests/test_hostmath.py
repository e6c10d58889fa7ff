"""CPU: the closed forms the CUDA kernels use (hs_phase.cuh, compiled here as plain C++ into a
test-only harness) against the dual-number oracle.  This is a check OF the kernel math, not a
CPU path of the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm():
    so = os.path.join(HERE, "hostmath", "libhostmath.so")
    src = os.path.join(HERE, "hostmath", "hostmath.cpp")
    hdrs = [os.path.join(HERE, "..", "hyperelasticsolver_b200", "csrc", h) for h in ("hs_phase.cuh", "hs_hank.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(f) for f in [src] + hdrs) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.hm_phase.argtypes = [dp, C.c_int, C.c_double, dp, C.c_double, dp, dp]
    lib.hm_sym3_eigs.argtypes = [dp, dp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("name,kw,gen", [("default", {}, 0), ("default-genericpath", {}, 1),
                                         ("alt", dict(c0=6.22, cv=9.0e-4, b0=3.16, beta=3.577, gamma=2.088), 1)])
def test_closed_forms_vs_dual_oracle(hm, oracle, name, kw, gen):
    rng = np.random.default_rng(1)
    eos = oracle.barton2009(**kw)
    rel = lambda a, b, s=None: np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() if s is None else s)
    for it in range(120):
        alpha = rng.uniform(0.05, 0.95)
        F = np.eye(3) + 0.1 * rng.uniform(-1, 1, (3, 3))
        u = rng.uniform(-2, 2, 3); S = rng.uniform(0, 2e-3)
        Pp = np.array([alpha, 8.9 / np.linalg.det(F), *u, S, *F.flatten(order="F")])
        Q, _ = oracle.prim2cons([eos, eos], 1, np.concatenate([Pp, Pp]))
        q = Q[:15]
        out = np.zeros(64)
        m = np.ascontiguousarray(q[2:5]); A = np.ascontiguousarray(q[6:15])
        hm.hm_phase(_p(eos), gen, q[0], _p(m), q[5], _p(A), _p(out))
        rho, uu, Sp, T, sig1, Gs, cmax, fl, S6, bad = out[0], out[1:4], out[5], out[6], out[7:10], out[10:16], out[16], out[17:32], out[32:38], out[38]
        assert bad == 0
        Po, _ = oracle.cons2prim([eos, eos], 1, Q); Po = Po[:15]
        Fo, So = Po[6:15], Po[5]
        Go = oracle.finger(Fo)
        assert rel(rho, Po[1]) < 1e-14 and rel(uu, Po[2:5], 1.0) < 1e-14
        assert rel(eos[2] * np.log(Sp), So, 1e-3) < 1e-12
        assert rel(T, oracle.temperature(eos, So, Go)) < 1e-12
        assert rel(Gs, Go[[0, 3, 6, 4, 7, 8]]) < 1e-13
        sig = oracle.stress(eos, So, Fo)
        assert rel(sig1, sig[[0, 3, 6]]) < 1e-12
        fo, _ = oracle.flux([eos, eos], 1, Q)
        assert rel(fl, fo[:15]) < 1e-12
        ac = oracle.acoustic(eos, So, Fo)
        assert rel(S6, [ac[0, 0], ac[0, 1], ac[0, 2], ac[1, 1], ac[1, 2], ac[2, 2]]) < 1e-12
        eg, _ = oracle.get_eigvals([eos, eos], 1, Q)
        assert rel(Po[2] + cmax, eg[0].max()) < 1e-13


def test_sym3_eigs(hm):
    """Jacobi (full get_eigvals API) is accurate for degenerate pairs; the trigonometric form (hot
    path) is accurate for the largest eigenvalue whenever it is simple."""
    dp = C.POINTER(C.c_double)
    hm.hm_sym3_eigs_jacobi.argtypes = [dp, dp]
    hm.hm_sym3_max_abs.argtypes = [dp]; hm.hm_sym3_max_abs.restype = C.c_double
    rng = np.random.default_rng(3)
    for it in range(300):
        M = rng.normal(size=(3, 3)); M = M + M.T
        if it % 10 == 0:  # degenerate smaller pair (shear speeds at F = I)
            Qm, _ = np.linalg.qr(rng.normal(size=(3, 3)))
            M = Qm @ np.diag([2.0, 2.0, 5.0]) @ Qm.T
        a = np.array([M[0, 0], M[0, 1], M[0, 2], M[1, 1], M[1, 2], M[2, 2]])
        ref = np.linalg.eigvalsh(M)
        scale = max(1.0, np.abs(ref).max())
        ev = np.zeros(3)
        hm.hm_sym3_eigs_jacobi(_p(a), _p(ev))
        assert np.abs(ev - ref).max() < 1e-14 * scale
        # largest |eigenvalue| from the trigonometric form: accurate when separated from its neighbour
        big = np.abs(ref).max()
        gap = min(abs(big - abs(x)) for x in ref if abs(abs(x) - big) > 0) if it % 10 else 3.0
        if gap > 1e-3 * scale:
            assert abs(hm.hm_sym3_max_abs(_p(a)) - big) < 2e-14 * scale
    a = np.array([3.0, 0, 0, 3.0, 0, 3.0]); ev = np.zeros(3)
    hm.hm_sym3_eigs_jacobi(_p(a), _p(ev))
    assert np.array_equal(ev, [3.0, 3.0, 3.0])
    assert hm.hm_sym3_max_abs(_p(a)) == 3.0


def test_acoustic_general_normal(hm, oracle):
    """closed form for an arbitrary unit normal (hs_get_eigvals) vs the nested-dual jacobian"""
    dp = C.POINTER(C.c_double)
    hm.hm_acoustic_n.argtypes = [dp, C.c_int, C.c_double, dp, C.c_double, dp, dp, dp]
    rng = np.random.default_rng(2)
    for eos, gen in ((oracle.barton2009(), 0), (oracle.barton2009(c0=6.22, cv=9e-4, b0=3.16, beta=3.577, gamma=2.088), 1)):
        for it in range(60):
            alpha = rng.uniform(0.1, 0.9); F = np.eye(3) + 0.1 * rng.uniform(-1, 1, (3, 3))
            u = rng.uniform(-1, 1, 3); S = rng.uniform(0, 1e-3)
            n = rng.normal(size=3); n /= np.linalg.norm(n)
            if it % 5 == 0:
                n = np.array([0.0, 1.0, 0.0])
            Pp = np.array([alpha, 8.9 / np.linalg.det(F), *u, S, *F.flatten(order="F")])
            Q, _ = oracle.prim2cons([eos, eos], 1, np.concatenate([Pp, Pp]))
            Po, _ = oracle.cons2prim([eos, eos], 1, Q)
            ac = oracle.acoustic(eos, Po[5], Po[6:15], n)
            S6 = np.zeros(6); m = np.ascontiguousarray(Q[2:5]); A = np.ascontiguousarray(Q[6:15]); nn = np.ascontiguousarray(n)
            hm.hm_acoustic_n(_p(eos), gen, Q[0], _p(m), Q[5], _p(A), _p(nn), _p(S6))
            ref = 0.5 * (ac + ac.T)
            r6 = np.array([ref[0, 0], ref[0, 1], ref[0, 2], ref[1, 1], ref[1, 2], ref[2, 2]])
            assert np.abs(S6 - r6).max() < 1e-12 * np.abs(r6).max()


def test_hank2016_closed_forms_vs_dual_oracle(hm, oracle):
    """hs_hank.cuh (what k_hank runs) against the literal dual-number restatement of EquationsOfState.jl:301-356"""
    dp = C.POINTER(C.c_double)
    hm.hm_hank_energy.argtypes = [dp, C.c_double, C.c_double, dp, dp]
    hm.hm_hank_pressure.argtypes = [dp, C.c_double, C.c_double, dp, dp]
    hm.hm_hank_stress.argtypes = [dp, C.c_double, dp, dp]
    rng = np.random.default_rng(7)
    for eos in (oracle.hank2016(), oracle.hank2016(rho0=8.9, mu=48e9, gamma=4.2, pres_inf=34e9, a=-0.3), oracle.hank2016(a=1.0)):
        for it in range(100):
            A = np.eye(3) + 0.2 * rng.uniform(-1, 1, (3, 3))
            a9 = np.ascontiguousarray(A.flatten(order="F"))
            den = eos[0] * abs(np.linalg.det(A)) * rng.uniform(0.9, 1.1)
            pres = rng.uniform(-1e9, 5e10)
            G = A.T @ A
            g9 = np.ascontiguousarray(G.flatten(order="F"))
            inv3 = oracle.invariants(g9)
            e_o, st = oracle.hank_energy(eos, den, pres, g9); assert st == 0
            out = np.zeros(1)
            assert hm.hm_hank_energy(_p(eos), den, pres, _p(g9), _p(out)) == 0
            assert abs(out[0] - e_o) <= 1e-13 * abs(e_o)
            p_o, st = oracle.hank_pressure(eos, den, e_o, inv3); assert st == 0
            assert hm.hm_hank_pressure(_p(eos), den, e_o, _p(inv3), _p(out)) == 0
            assert abs(out[0] - p_o) <= 1e-12 * max(abs(p_o), eos[2] * eos[3])
            assert abs(p_o - pres) <= 1e-12 * eos[2] * eos[3]            # pressure inverts energy
            s_o, st = oracle.hank_stress(eos, den, pres, a9); assert st == 0
            sig = np.zeros(9)
            assert hm.hm_hank_stress(_p(eos), den, _p(a9), _p(sig)) == 0
            assert np.abs(sig - s_o).max() <= 1e-12 * max(np.abs(s_o).max(), 1e-3 * eos[1])
    # domain: det G <= 0 is where Julia's fractional power throws
    bad = np.diag([1.0, 1.0, -1.0]).flatten()
    out = np.zeros(1)
    assert hm.hm_hank_energy(_p(eos), 2.7, 1e9, _p(bad), _p(out)) == 1
    assert oracle.hank_energy(eos, 2.7, 1e9, bad)[1] == 1


def test_row1_quadrature_state_matches_full_state(hm, oracle):
    """phase_state_row1 (B = A A^T + Cayley-Hamilton; tuning build HS_PHASE_CH) against the full phase_state: everything the
    non-conservative column reads (u, T, row 1 of sigma) to <= 5e-13, inside the 1e-12 per-evaluation budget."""
    dp = C.POINTER(C.c_double)
    hm.hm_phase_row1.argtypes = [dp, C.c_int, C.c_double, dp, C.c_double, dp, dp]
    rng = np.random.default_rng(4)
    worst = 0.0
    for eos, gen in ((oracle.barton2009(), 0), (oracle.barton2009(), 1), (oracle.barton2009(c0=6.22, cv=9e-4, b0=3.16, beta=3.577, gamma=2.088), 1)):
        for it in range(200):
            alpha = rng.uniform(0.05, 0.95)
            F = np.eye(3) + (0.3 if it % 4 == 0 else 0.1) * rng.uniform(-1, 1, (3, 3))
            u = rng.uniform(-2, 2, 3); S = rng.uniform(0, 2e-3)
            Pp = np.array([alpha, 8.9 / np.linalg.det(F), *u, S, *F.flatten(order="F")])
            Q, _ = oracle.prim2cons([eos, eos], 1, np.concatenate([Pp, Pp]))
            q = Q[:15]
            m = np.ascontiguousarray(q[2:5]); A = np.ascontiguousarray(q[6:15])
            full = np.zeros(64); r1 = np.zeros(32)
            hm.hm_phase(_p(eos), gen, q[0], _p(m), q[5], _p(A), _p(full))
            hm.hm_phase_row1(_p(eos), gen, q[0], _p(m), q[5], _p(A), _p(r1))
            assert r1[14] == 0 and full[38] == 0
            assert np.array_equal(r1[0:5], full[0:5])                       # rho, u, Etot: same expressions
            sig_scale = max(np.abs(full[7:10]).max(), 1.0)
            e_T = abs(r1[5] - full[6]) / abs(full[6]); e_s = np.abs(r1[6:9] - full[7:10]).max() / sig_scale
            worst = max(worst, e_T, e_s)
            assert e_T < 5e-13 and e_s < 5e-13, (e_T, e_s)   # (T carries the cancellation of e - W - U_cold: the full state holds 1e-12 vs the oracle)
            # against the dual-number oracle as well
            Po, _ = oracle.cons2prim([eos, eos], 1, Q); Po = Po[:15]
            sig = oracle.stress(eos, Po[5], Po[6:15])
            assert np.abs(r1[6:9] - sig[[0, 3, 6]]).max() < 1e-12 * max(np.abs(sig).max(), 1.0)
            To = oracle.temperature(eos, Po[5], oracle.finger(Po[6:15]))
            assert abs(r1[5] - To) < 1e-12 * abs(To)
    print("row1 worst relative deviation", worst)


def test_entropy_clamp_branch(hm, oracle):
    """EquationsOfState.jl:152-154: e_int below the cold curve -> S' = 1e-6 exactly.  The kernels select
    S' and T = t0 I3^(gamma/2) S' in that branch instead of forming them through 1 + th/(cv t0 I3^(gamma/2)), which cancels
    six digits.  Also low-entropy UNclamped states: there S' = 1 + th_raw/(...) is ill-conditioned in the reference itself
    (the reference's own roundoff in e_int = E - |u|^2/2 is amplified by 1/(cv T)), so agreement is asserted at that
    conditioning bound, not at 1e-12."""
    rng = np.random.default_rng(7)
    eos = oracle.barton2009()
    rel = lambda a, b, s=None: np.abs(np.asarray(a) - np.asarray(b)).max() / (np.abs(np.asarray(b)).max() if s is None else s)
    n_clamped = 0
    for it in range(200):
        alpha = rng.uniform(0.05, 0.95)
        F = np.eye(3) + 0.1 * rng.uniform(-1, 1, (3, 3))
        u = rng.uniform(-2, 2, 3); S = rng.uniform(0, 4e-4)
        Pp = np.array([alpha, 8.9 / np.linalg.det(F), *u, S, *F.flatten(order="F")])
        Q, _ = oracle.prim2cons([eos, eos], 1, np.concatenate([Pp, Pp]))
        Q = Q.copy()
        Q[5] -= Q[1] * rng.uniform(0.4, 2.0)       # push the internal energy below the cold curve
        Q[20] = Q[5] * Q[15] / Q[0]                # (second phase: same state, its own alpha)
        assert np.isclose(Q[1] / Q[0], Q[16] / Q[15])
        q = Q[:15]
        out = np.zeros(64)
        m = np.ascontiguousarray(q[2:5]); A = np.ascontiguousarray(q[6:15])
        hm.hm_phase(_p(eos), 0, q[0], _p(m), q[5], _p(A), _p(out))
        Sp, T, sig1, cmax, fl, S6, bad = out[5], out[6], out[7:10], out[16], out[17:32], out[32:38], out[38]
        Po, _ = oracle.cons2prim([eos, eos], 1, Q); Po = Po[:15]
        Fo, So = Po[6:15], Po[5]
        Go = oracle.finger(Fo)
        if not np.isclose(So, eos[2] * np.log(1e-6), rtol=1e-12):
            continue
        n_clamped += 1
        assert bad == 0
        assert Sp == 1e-6                                           # selected, not computed
        assert rel(T, oracle.temperature(eos, So, Go)) < 1e-13
        assert rel(sig1, oracle.stress(eos, So, Fo)[[0, 3, 6]]) < 1e-12
        fo, _ = oracle.flux([eos, eos], 1, Q)
        assert rel(fl, fo[:15]) < 1e-12
        ac = oracle.acoustic(eos, So, Fo)
        assert rel(S6, [ac[0, 0], ac[0, 1], ac[0, 2], ac[1, 1], ac[1, 2], ac[2, 2]]) < 1e-12
        eg, _ = oracle.get_eigvals([eos, eos], 1, Q)
        assert rel(Po[2] + cmax, eg[0].max()) < 1e-13
    assert n_clamped >= 150
    # low-entropy states that are NOT clamped (S' between 1e-5 and 1e-2)
    worst = 0.0
    for it in range(200):
        alpha = rng.uniform(0.05, 0.95)
        F = np.eye(3) + 0.1 * rng.uniform(-1, 1, (3, 3))
        u = rng.uniform(-2, 2, 3)
        Sprime = 10.0 ** rng.uniform(-5, -2)
        Pp = np.array([alpha, 8.9 / np.linalg.det(F), *u, eos[2] * np.log(Sprime), *F.flatten(order="F")])
        Q, _ = oracle.prim2cons([eos, eos], 1, np.concatenate([Pp, Pp]))
        q = Q[:15]
        out = np.zeros(64)
        m = np.ascontiguousarray(q[2:5]); A = np.ascontiguousarray(q[6:15])
        hm.hm_phase(_p(eos), 0, q[0], _p(m), q[5], _p(A), _p(out))
        Po, _ = oracle.cons2prim([eos, eos], 1, Q); Po = Po[:15]
        Go = oracle.finger(Po[6:15])
        To = oracle.temperature(eos, Po[5], Go)
        # T = (e_int - W - U_cold)/cv + t0 I3^(gamma/2): roundoff eps (|E| + |u|^2/2) of e_int = E - |u|^2/2 shows up as
        # eps (|E| + |u|^2/2)/(cv T) relative in T -- in the reference's own arithmetic as much as in the kernels'
        cond = (abs(q[5] / q[1]) + 0.5 * (Po[2:5] ** 2).sum()) / (eos[2] * To)
        worst = max(worst, abs(out[6] - To) / To / (cond * 2.2e-16))
        assert abs(out[6] - To) / To < 64 * 2.2e-16 * max(cond, 1.0)
        # everything that does not divide by S' stays at 1e-12
        assert rel(out[7:10], oracle.stress(eos, Po[5], Po[6:15])[[0, 3, 6]]) < 1e-12
    assert worst < 64, worst
    print('worst T error in units of eps * cond:', worst)
