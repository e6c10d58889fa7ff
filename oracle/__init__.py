"""CPU parity oracle -- TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_build/liboracle.so (built from oracle.cpp by `make -C oracle`).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  The product package (hyperelasticsolver_b200) never does.

PARITY UNPINNED (see oracle.cpp header and DESIGN.md): the reference ships no golden vectors
and Julia is not installed, so the oracle is pinned against SURVEY.md Appendix B, an
independent torch-autograd restatement (oracle/pyoracle.py) and physical anchors only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

SP13, MPH30 = 0, 1
LXF, HLL = 0, 1
NVAR = {SP13: 13, MPH30: 30}
NEIG = {SP13: 6, MPH30: 12}


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("oracle.cpp", "dual.hpp")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None
_dp = C.POINTER(C.c_double)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        L = _lib
        i64 = C.c_int64
        L.hso_cons2prim.argtypes = [_dp, C.c_int, _dp, _dp, i64]
        L.hso_prim2cons.argtypes = [_dp, C.c_int, _dp, _dp, i64]
        L.hso_flux.argtypes = [_dp, C.c_int, _dp, _dp, i64]
        L.hso_noncons_flux.argtypes = [_dp, _dp, _dp, _dp, i64]
        L.hso_get_eigvals.argtypes = [_dp, C.c_int, _dp, _dp, i64]
        L.hso_get_eigvals_n.argtypes = [_dp, _dp, _dp, _dp, i64]
        L.hso_energy.argtypes = [_dp, C.c_double, _dp]; L.hso_energy.restype = C.c_double
        L.hso_entropy.argtypes = [_dp, C.c_double, _dp]; L.hso_entropy.restype = C.c_double
        L.hso_temperature.argtypes = [_dp, C.c_double, _dp]; L.hso_temperature.restype = C.c_double
        L.hso_finger.argtypes = [_dp, _dp]
        L.hso_hank_energy.argtypes = [_dp, C.c_double, C.c_double, _dp, _dp]
        L.hso_hank_pressure.argtypes = [_dp, C.c_double, C.c_double, _dp, _dp]
        L.hso_hank_stress.argtypes = [_dp, C.c_double, C.c_double, _dp, _dp]
        L.hso_invariants.argtypes = [_dp, _dp]
        L.hso_stress.argtypes = [_dp, C.c_double, _dp, _dp]
        L.hso_acoustic.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.hso_quadrature.argtypes = [C.c_int, _dp, _dp]
        L.hso_hll.argtypes = [_dp] * 9 + [i64]
        L.hso_lxf.argtypes = [_dp, _dp, _dp, C.c_double, _dp, _dp, _dp, i64]
        L.hso_sp_hll.argtypes = [_dp] * 6 + [i64]
        L.hso_run.argtypes = [_dp, C.c_int, C.c_int, _dp, i64, i64, C.c_double, C.c_double, C.c_double, i64,
                              _dp, C.POINTER(i64), _dp, C.c_int, C.c_int]
        L.hso_lambda_max.argtypes = [_dp, C.c_int, _dp, i64, _dp]; L.hso_lambda_max.restype = C.c_double
        L.hso_hardware_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def barton2009(rho0=8.93, c0=4.6, cv=3.9e-4, t0=300.0, b0=2.1, alpha=1.0, beta=3.0, gamma=2.0):
    """EquationsOfState.jl:90-115 -> the 10-double parameter block."""
    b0sq = b0 ** 2
    k0 = c0 ** 2 - (4 / 3) * b0 ** 2
    return np.array([rho0, c0, cv, t0, b0, alpha, beta, gamma, b0sq, k0], dtype=np.float64)


def eos_block(eos=None, nphase=2):
    if eos is None:
        eos = [barton2009()] * nphase
    e = _f(np.stack([_f(x) for x in eos]))
    assert e.shape[1] == 10
    if e.shape[0] == 1:  # SP13 entry points only read the first block
        e = _f(np.concatenate([e, e]))
    return e


# All arrays are "Julia layout": shape (n, nvar) C-contiguous == (nvar, n) column-major.
def cons2prim(eos, model, Q):
    Q = _f(Q); P = np.empty_like(Q)
    st = lib().hso_cons2prim(_p(eos_block(eos)), model, _p(Q), _p(P), Q.size // NVAR[model])
    return P, st


def prim2cons(eos, model, P):
    P = _f(P); Q = np.empty_like(P)
    st = lib().hso_prim2cons(_p(eos_block(eos)), model, _p(P), _p(Q), P.size // NVAR[model])
    return Q, st


def flux(eos, model, Q):
    Q = _f(Q); F = np.empty_like(Q)
    st = lib().hso_flux(_p(eos_block(eos)), model, _p(Q), _p(F), Q.size // NVAR[model])
    return F, st


def noncons_cols(eos, Q):
    Q = _f(Q); col = np.empty_like(Q)
    st = lib().hso_noncons_flux(_p(eos_block(eos)), _p(Q), _p(col), None, Q.size // 30)
    return col, st


def noncons_dense(eos, Q):
    Q = _f(Q).reshape(-1, 30); B = np.empty((Q.shape[0], 30, 30))
    lib().hso_noncons_flux(_p(eos_block(eos)), _p(Q), None, _p(B), Q.shape[0])
    return B.transpose(0, 2, 1)  # [n, row, col]


def get_eigvals(eos, model, Q):
    Q = _f(Q); n = Q.size // NVAR[model]
    eig = np.empty((n, NEIG[model]))
    st = lib().hso_get_eigvals(_p(eos_block(eos)), model, _p(Q), _p(eig), n)
    return eig, st


def get_eigvals_n(eos, Q, n):
    """two-phase get_eigvals for an arbitrary normal n"""
    Q = _f(Q); cnt = Q.size // 30
    eig = np.empty((cnt, 12))
    st = lib().hso_get_eigvals_n(_p(eos_block(eos)), _p(Q), _p(_f(n)), _p(eig), cnt)
    return eig, st


def energy(eos1, S, G): return lib().hso_energy(_p(_f(eos1)), float(S), _p(_f(G)))
def entropy(eos1, e, G): return lib().hso_entropy(_p(_f(eos1)), float(e), _p(_f(G)))
def temperature(eos1, S, G): return lib().hso_temperature(_p(_f(eos1)), float(S), _p(_f(G)))


def hank2016(rho0=2.7, mu=26e9, gamma=3.4, pres_inf=21.5e9, a=0.5):
    """EquationsOfState.jl:312-319 -> the 5-double parameter block."""
    return np.array([rho0, mu, gamma, pres_inf, a], dtype=np.float64)


def hank_energy(eos, den, pres, G):
    """EquationsOfState.jl:317-331; G: 9 column-major entries.  Returns (e_int, domain_error)."""
    e = np.empty(1); st = lib().hso_hank_energy(_p(_f(eos)), float(den), float(pres), _p(_f(G)), _p(e)); return e[0], st


def hank_pressure(eos, den, e_int, inv3):
    """EquationsOfState.jl:333-346; inv3 = invariants(G)."""
    e = np.empty(1); st = lib().hso_hank_pressure(_p(_f(eos)), float(den), float(e_int), _p(_f(inv3)), _p(e)); return e[0], st


def hank_stress(eos, den, pres, A):
    """EquationsOfState.jl:348-356; A (distortion) and the result: 9 column-major entries."""
    s = np.empty(9); st = lib().hso_hank_stress(_p(_f(eos)), float(den), float(pres), _p(_f(A)), _p(s)); return s, st


def finger(F):
    G = np.empty(9); lib().hso_finger(_p(_f(F)), _p(G)); return G


def invariants(G):
    i = np.empty(3); lib().hso_invariants(_p(_f(G)), _p(i)); return i


def stress(eos1, S, F):
    s = np.empty(9); lib().hso_stress(_p(_f(eos1)), float(S), _p(_f(F)), _p(s)); return s


def acoustic(eos1, S, F, n=(1.0, 0.0, 0.0)):
    a = np.empty(9); lib().hso_acoustic(_p(_f(eos1)), float(S), _p(_f(F)), _p(_f(n)), _p(a))
    return a.reshape(3, 3).T  # [i, j]


def quadrature(lobatto=False):
    x = np.empty(6); w = np.empty(6); lib().hso_quadrature(int(lobatto), _p(x), _p(w)); return x, w


def hll(eos, Ql, Qr, eig_l, eig_r):
    Ql = _f(Ql).reshape(-1, 30); Qr = _f(Qr).reshape(-1, 30); n = Ql.shape[0]
    eig_l = _f(eig_l).reshape(n, 12); eig_r = _f(eig_r).reshape(n, 12)
    cons = np.empty_like(Ql); dm = np.empty_like(Ql); dp = np.empty_like(Ql); s = np.empty((n, 2))
    st = lib().hso_hll(_p(eos_block(eos)), _p(Ql), _p(Qr), _p(eig_l), _p(eig_r), _p(cons), _p(dm), _p(dp), _p(s), n)
    return cons, dm, dp, s, st


def lxf(eos, Ql, Qr, lam):
    Ql = _f(Ql).reshape(-1, 30); Qr = _f(Qr).reshape(-1, 30); n = Ql.shape[0]
    cons = np.empty_like(Ql); dm = np.empty_like(Ql); dp = np.empty_like(Ql)
    st = lib().hso_lxf(_p(eos_block(eos)), _p(Ql), _p(Qr), float(lam), _p(cons), _p(dm), _p(dp), n)
    return cons, dm, dp, st


def sp_hll(eos, Ql, Qr, eig_l, eig_r):
    Ql = _f(Ql).reshape(-1, 13); Qr = _f(Qr).reshape(-1, 13); n = Ql.shape[0]
    cons = np.empty_like(Ql)
    st = lib().hso_sp_hll(_p(eos_block(eos)), _p(Ql), _p(Qr), _p(_f(eig_l)), _p(_f(eig_r)), _p(cons), n)
    return cons, st


def lambda_max(eos, model, Q):
    Q = _f(Q); n = Q.size // NVAR[model]
    return lib().hso_lambda_max(_p(eos_block(eos)), model, _p(Q), n, None)


def run(eos, model, fluxkind, Q, cfl, dx, t_end, max_steps, nthreads=1, literal=False, t0=0.0):
    """main.jl:202-227.  Q: (nprob, ncells, nvar) or (ncells, nvar).  Returns dict."""
    Q = _f(Q).copy()
    nvar = NVAR[model]
    shp = Q.shape
    Q3 = Q.reshape(-1, shp[-2], nvar)
    nprob, ncells = Q3.shape[0], Q3.shape[1]
    t = np.full(nprob, float(t0)); steps = np.zeros(nprob, dtype=np.int64)
    hist = np.zeros((nprob, max_steps))
    st = lib().hso_run(_p(eos_block(eos)), model, fluxkind, _p(Q3), ncells, nprob, cfl, dx, t_end, max_steps,
                       _p(t), steps.ctypes.data_as(C.POINTER(C.c_int64)), _p(hist), nthreads, int(literal))
    return dict(Q=Q3.reshape(shp), t=t, steps=steps, dt=hist, status=st)


def hardware_threads():
    return lib().hso_hardware_threads()
