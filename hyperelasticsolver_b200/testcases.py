"""Primitive left/right Riemann states of the reference's shipped test cases.

Host-side setup data only (no physics): mirrors the tables of
`initial_states`, HyperelasticityMPh.jl:275-420 (two-phase, 30 primitives
`[alpha, rho, u(3), S, F(9 column-major)]` per phase) and Hyperelasticity.jl:124-165
(single-phase, `[u(3), F(9 row-major), S]`).  The conservative states are obtained by passing
these through `prim2cons_mph` / `prim2cons` of the library (HyperelasticityMPh.jl:422-423).
"""
from __future__ import annotations

import numpy as np

_I = [[1, 0, 0], [0, 1, 0], [0, 0, 1]]
_F3 = [[1.0, 0.0, 0.0], [-0.01, 0.95, 0.02], [-0.015, 0.0, 0.9]]
_F4L = [[0.98, 0.0, 0.0], [0.02, 1.0, 0.1], [0.0, 0.0, 1.0]]
_F4R = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.1], [0.0, 0.0, 1.0]]
_F5R = [[1.0, 0.0, 0.0], [0.015, 0.95, 0.0], [-0.01, 0.0, 0.9]]

# testcase -> (alpha_l_1, alpha_l_2, alpha_r_1, alpha_r_2, den, u_l, S_l, F_l, u_r, S_r, F_r)
# alpha literals are the reference's own (0.1 is a literal, not 1-0.9: HyperelasticityMPh.jl:349-352).
MPH_CASES = {
    1: (0.5, 0.5, 0.5, 0.5, 5.0, [0, 0, 0], 0, _I, [0, 0, 0], 0, _I),
    2: (0.5, 0.5, 0.5, 0.5, 5.0, [1.0, 0, 0], 0, _I, [1.0, 0, 0], 0, _I),
    3: (0.5, 0.5, 0.5, 0.5, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [2.0, 0.0, 0.1], 0.0, _F3),
    4: (0.5, 0.5, 0.5, 0.5, 8.9, [0.0, 0.5, 1.0], 1e-3, _F4L, [0.0, 0.0, 0.0], 0.0, _F4R),
    5: (0.5, 0.5, 0.5, 0.5, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [0.0, -0.03, -0.01], 0.0, _F5R),
    6: (0.1, 0.9, 0.9, 0.1, 8.9, [0.0, 0.5, 1.0], 1.0e-3, _F4L, [0.0, 0.0, 0.0], 0.0, _F4R),
    7: (0.1, 0.9, 0.9, 0.1, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [0.0, -0.03, -0.01], 0.0, _F5R),
    10: (0.4, 0.6, 0.6, 0.4, 8.9, [2.0, 0.0, 0.1], 0.0, _F3, [2.0, 0.0, 0.1], 0.0, _F3),
}


def _det3(F):
    F = np.asarray(F, dtype=np.float64)
    return float(np.linalg.det(F))  # LinearAlgebra.det == LAPACK LU, HyperelasticityMPh.jl:412-415


def mph_primitive_states(testcase: int):
    """(Pl, Pr), each 30 primitives, as assembled at HyperelasticityMPh.jl:412-420."""
    if testcase not in MPH_CASES:
        raise ValueError(f"unknown multiphase test case {testcase}")
    a_l1, a_l2, a_r1, a_r2, den, u_l, S_l, F_l, u_r, S_r, F_r = MPH_CASES[testcase]
    F_l = np.asarray(F_l, dtype=np.float64)
    F_r = np.asarray(F_r, dtype=np.float64)
    den_l = den / _det3(F_l)
    den_r = den / _det3(F_r)

    def phase(a, d, u, S, F):
        return [a, d, *map(float, u), float(S), *F.flatten(order="F")]  # F... splats column-major

    Pl = np.array(phase(a_l1, den_l, u_l, S_l, F_l) + phase(a_l2, den_l, u_l, S_l, F_l))
    Pr = np.array(phase(a_r1, den_r, u_r, S_r, F_r) + phase(a_r2, den_r, u_r, S_r, F_r))
    return Pl, Pr


_R3 = 3 ** 0.5
SP_CASES = {  # Hyperelasticity.jl:125-160
    1: ([0.0, 0.5, 1.0], _F4L, 1e-3, [0.0, 0.0, 0.0], _F4R, 0.0),
    2: ([2.0, 0.0, 0.1], _F3, 0.0, [0.0, -0.03, -0.01], _F5R, 0.0),
    3: ([1.0, 0.0, 0.0], [[0.5, -0.5 * _R3, 0.0], [0.5 * _R3, 0.5, 0.0], [0.0, 0.0, 1.0]], 0.0,
        [1.0, 0.0, 0.0], [[0.5, -0.5 * _R3, 0.0], [0.5 * _R3, 0.5, 0.0], [0.0, 0.0, 1.0]], 0.0),
}


def sp_primitive_states(testcase: int):
    """(Pl, Pr), each `[u(3), F(9 row-major), S]` = the arguments of Hyperelasticity.jl:70."""
    if testcase in SP_CASES:
        u_l, F_l, S_l, u_r, F_r, S_r = SP_CASES[testcase]
    else:  # Hyperelasticity.jl:161-165
        u_l = u_r = [0.0, 0.0, 0.0]; F_l = F_r = _I; S_l = S_r = 0.0
    mk = lambda u, F, S: np.array([*map(float, u), *np.asarray(F, dtype=np.float64).flatten(order="C"), float(S)])
    return mk(u_l, F_l, S_l), mk(u_r, F_r, S_r)


def riemann_grid(Ql, Qr, nx: int):
    """initial_condition, main.jl:99-106: cell i (1-based) is Ql iff (i-1) < nx/2.  -> (nx, nvar)."""
    Ql = np.asarray(Ql, dtype=np.float64); Qr = np.asarray(Qr, dtype=np.float64)
    i0 = np.arange(nx)
    return np.where((i0 < nx / 2)[:, None], Ql[None, :], Qr[None, :]).copy()
