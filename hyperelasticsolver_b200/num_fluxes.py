"""Mirror of the reference module `NumFluxes` (NumFluxes.jl:15 exports `lxf, hll`), batched over
faces.  Both return the reference's triple `(cons, noncons_minus, noncons_plus)`."""
from __future__ import annotations

import numpy as np

from . import _lib as L

__all__ = ["lxf", "hll"]


def _model_of(Q):
    nvar = np.shape(Q)[-1]
    if nvar == 30:
        return L.MPH30
    if nvar == 13:
        return L.SP13
    raise ValueError(f"state vectors must have 30 (two-phase) or 13 (single-phase) entries, got {nvar}")


def hll(eos, Q_l, Q_r, eigvals, device=0, return_speeds=False):
    """NumFluxes.jl:70-132.  `eigvals = [eig_l, eig_r]`: get_eigvals of the two cells, as passed by
    update_cell (main.jl:56-57).  Returns (zeros, D^-, D^+) for the two-phase model
    (NumFluxes.jl:82); for the single-phase model (cons, 0, 0) with the conservative HLL flux of
    NumFluxes.jl:75-78."""
    model = _model_of(Q_l)
    nvar, neig = L.NVAR[model], 6 * L.NPHASE[model]
    ql = np.ascontiguousarray(Q_l, dtype=np.float64)
    qr = np.ascontiguousarray(Q_r, dtype=np.float64)
    el = np.ascontiguousarray(eigvals[0], dtype=np.float64)
    er = np.ascontiguousarray(eigvals[1], dtype=np.float64)
    n = ql.size // nvar
    if qr.shape != ql.shape or el.size != n * neig or er.size != n * neig:
        raise ValueError("inconsistent shapes")
    cons = np.empty_like(ql); dm = np.empty_like(ql); dp = np.empty_like(ql)
    s = np.empty(ql.shape[:-1] + (2,))
    L.check(L.lib().hs_hll(model, L.eos_array(eos, model), L.NPHASE[model], ql.ctypes.data, qr.ctypes.data, el.ctypes.data,
                           er.ctypes.data, cons.ctypes.data, dm.ctypes.data, dp.ctypes.data, s.ctypes.data, n, device))
    return (cons, dm, dp, s) if return_speeds else (cons, dm, dp)


def lxf(eos, Q_l, Q_r, lam, device=0):
    """NumFluxes.jl:25-60; `lam` is dx/dt."""
    model = _model_of(Q_l)
    nvar = L.NVAR[model]
    ql = np.ascontiguousarray(Q_l, dtype=np.float64)
    qr = np.ascontiguousarray(Q_r, dtype=np.float64)
    n = ql.size // nvar
    cons = np.empty_like(ql); dm = np.empty_like(ql); dp = np.empty_like(ql)
    L.check(L.lib().hs_lxf(model, L.eos_array(eos, model), L.NPHASE[model], ql.ctypes.data, qr.ctypes.data, float(lam),
                           cons.ctypes.data, dm.ctypes.data, dp.ctypes.data, n, device))
    return cons, dm, dp
