"""Mirror of the reference module `HyperelasticityMPh` (HyperelasticityMPh.jl:13 exports
`prim2cons_mph, cons2prim_mph, flux_mph, noncons_flux, initial_states, get_eigvals`), batched
and executed on the GPU through the C ABI.

Array convention: a single state is a length-30 vector; a batch is (n, 30) C-contiguous, which
is byte-identical to Julia's `Array{Float64,2}(30, n)`.  `eos` is a pair of `Barton2009`
(`eos::Tuple{T,T}`, main.jl:134).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .testcases import mph_primitive_states

__all__ = ["prim2cons_mph", "cons2prim_mph", "flux_mph", "noncons_flux", "initial_states", "get_eigvals"]

_MODEL = L.MPH30


def _batch(x, nvar):
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.ndim == 0 or a.shape[-1] != nvar:
        raise ValueError(f"expected trailing dimension {nvar}, got shape {a.shape}")
    return a, a.shape, a.size // nvar


def _cellop(fn, eos, X, nvar_out=None, device=0):
    a, shp, n = _batch(X, 30)
    out = np.empty_like(a)
    L.check(fn(_MODEL, L.eos_array(eos, _MODEL), 2, a.ctypes.data, out.ctypes.data, n, device))
    return out.reshape(shp)


def prim2cons_mph(eos, P, device=0):
    """HyperelasticityMPh.jl:63-90."""
    return _cellop(L.lib().hs_prim2cons, eos, P, device=device)


def cons2prim_mph(eos, Q, device=0):
    """HyperelasticityMPh.jl:99-133 (rho is recomputed from det(Q_F/alpha), :113-114)."""
    return _cellop(L.lib().hs_cons2prim, eos, Q, device=device)


def flux_mph(eos, Q, device=0):
    """HyperelasticityMPh.jl:140-175."""
    return _cellop(L.lib().hs_flux, eos, Q, device=device)


def noncons_flux(eos, Q, dense=True, device=0):
    """HyperelasticityMPh.jl:178-250.  dense=True returns the 30x30 matrix B (n, 30, 30) [row, col]
    exactly as the reference does; dense=False returns only its non-zero columns (n, 30)."""
    a, shp, n = _batch(Q, 30)
    col = np.empty_like(a)
    B = np.empty((n, 30, 30)) if dense else None
    L.check(L.lib().hs_noncons_flux(L.eos_array(eos, _MODEL), a.ctypes.data, col.ctypes.data,
                                    B.ctypes.data if dense else None, n, device))
    if not dense:
        return col.reshape(shp)
    B = B.transpose(0, 2, 1)  # library writes column-major (30,30,n) like Julia
    return B[0] if a.ndim == 1 else B.reshape(shp[:-1] + (30, 30))


def get_eigvals(eos, Q, n=(1, 0, 0), device=0):
    """HyperelasticityMPh.jl:252-266: per phase [u.n + c_k, u.n - c_k], k = 1..3, for a unit normal `n`
    (the reference's 1-D driver passes [1, 0, 0], main.jl:208)."""
    nn = np.ascontiguousarray(n, dtype=np.float64)
    if nn.shape != (3,):
        raise ValueError("n must have 3 components")
    a, shp, cnt = _batch(Q, 30)
    eig = np.empty(shp[:-1] + (12,))
    L.check(L.lib().hs_get_eigvals(_MODEL, L.eos_array(eos, _MODEL), 2, a.ctypes.data, nn.ctypes.data, eig.ctypes.data, cnt, device))
    return eig


def initial_states(eos, testcase: int, device=0):
    """HyperelasticityMPh.jl:275-426: (Ql, Qr) of the shipped Riemann test cases."""
    Pl, Pr = mph_primitive_states(testcase)
    Q = prim2cons_mph(eos, np.stack([Pl, Pr]), device=device)
    return Q[0].copy(), Q[1].copy()
