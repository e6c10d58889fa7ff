"""Single-phase (13-variable) model: the specification in `Hyperelasticity.jl` (stale in the
reference) made runnable with the shipped EoS variants.

Conservative vector  Q = [rho*u (3), rho*F (9, row-major), rho*E]      (Hyperelasticity.jl:81-91)
Primitive vector     P = [u (3), F (9, row-major), S]  = the arguments of prim2cons (:70).
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .testcases import sp_primitive_states

__all__ = ["prim2cons", "cons2prim", "flux", "get_eigvals", "initial_states"]

_MODEL = L.SP13


def _op(fn, eos, X, device):
    a = np.ascontiguousarray(X, dtype=np.float64)
    if a.shape[-1] != 13:
        raise ValueError(f"expected trailing dimension 13, got {a.shape}")
    out = np.empty_like(a)
    L.check(fn(_MODEL, L.eos_array(eos, _MODEL), 1, a.ctypes.data, out.ctypes.data, a.size // 13, device))
    return out


def prim2cons(eos, P, device=0):
    """Hyperelasticity.jl:70-93 (rho = rho0/det F)."""
    return _op(L.lib().hs_prim2cons, eos, P, device)


def cons2prim(eos, Q, device=0):
    """Hyperelasticity.jl:25-36 with density(), EquationsOfState.jl:259-264."""
    return _op(L.lib().hs_cons2prim, eos, Q, device)


def flux(eos, Q, device=0):
    """Hyperelasticity.jl:99-114."""
    return _op(L.lib().hs_flux, eos, Q, device)


def get_eigvals(eos, Q, n=(1, 0, 0), device=0):
    a = np.ascontiguousarray(Q, dtype=np.float64)
    nn = np.ascontiguousarray(n, dtype=np.float64)
    eig = np.empty(a.shape[:-1] + (6,))
    L.check(L.lib().hs_get_eigvals(_MODEL, L.eos_array(eos, _MODEL), 1, a.ctypes.data, nn.ctypes.data, eig.ctypes.data, a.size // 13, device))
    return eig


def initial_states(eos, testcase: int, device=0):
    """Hyperelasticity.jl:124-172."""
    Pl, Pr = sp_primitive_states(testcase)
    Q = prim2cons(eos, np.stack([Pl, Pr]), device=device)
    return Q[0].copy(), Q[1].copy()
