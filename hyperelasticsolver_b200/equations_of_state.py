"""Mirror of the `Hank2016` part of the reference module `EquationsOfState` (EquationsOfState.jl:301-366 exports
`Hank2016, eos_hank2016, pressure`; `energy` and `stress` are the module's generic functions, :12), batched and
executed on the GPU through the C ABI (`hs_hank2016_*`).

In the reference this material law is dead code and two of its three functions cannot run as written (a 3x3 Matrix
is handed to the Vector-only `invariants` / `finger`); here a 3x3 tensor is a (3, 3) array in Julia's column-major
reading or its 9 column-major entries -- see `csrc/hs_hank.cuh`.  The Barton2009 methods of `energy / entropy /
stress / acoustic` are reached through `cons2prim_mph`, `flux_mph` and `get_eigvals`, as in the reference's driver.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from ._lib import Hank2016

__all__ = ["Hank2016", "eos_hank2016", "energy", "pressure", "stress"]

eos_hank2016 = Hank2016()   # EquationsOfState.jl:364


def _scal(x):
    a = np.ascontiguousarray(x, dtype=np.float64)
    return a.reshape(-1), a.ndim == 0


def _tensor(x, nt):
    """(3, 3) Julia matrix (given as numpy [row, col]) -> its 9 column-major entries; (..., nt) batches pass through."""
    a = np.asarray(x, dtype=np.float64)
    if nt == 9 and a.shape == (3, 3):
        a = a.T
    a = np.ascontiguousarray(a).reshape(-1, nt)
    return a


def _call(fn, eos, s0, s1, ten, nt, no, device):
    if not isinstance(eos, Hank2016):
        raise TypeError("eos must be a Hank2016")
    a0, scalar = _scal(s0)
    a1, _ = _scal(s1)
    t = _tensor(ten, nt)
    n = t.shape[0]
    if a0.size != n or a1.size != n:
        raise ValueError(f"{n} tensors but {a0.size} / {a1.size} scalars")
    out = np.empty((n, no))
    L.check(fn(C.byref(eos), a0.ctypes.data, a1.ctypes.data, t.ctypes.data, out.ctypes.data, n, device))
    return out, scalar


def energy(eos, den, pres, G, device=0):
    """energy(eos::Hank2016, den, pres, G), EquationsOfState.jl:317-331."""
    out, scalar = _call(L.lib().hs_hank2016_energy, eos, den, pres, G, 9, 1, device)
    return float(out[0, 0]) if scalar else out[:, 0]


def pressure(eos, den, e_int, i, device=0):
    """pressure(eos::Hank2016, den, e_int, i), EquationsOfState.jl:333-346; i = invariants(G) (Strains.jl:46-52)."""
    out, scalar = _call(L.lib().hs_hank2016_pressure, eos, den, e_int, i, 3, 1, device)
    return float(out[0, 0]) if scalar else out[:, 0]


def stress(eos, den, pressure, distortion, device=0):
    """stress(eos::Hank2016, den, pressure, distortion), EquationsOfState.jl:348-356: 9 column-major entries of
    -2 den G de/dG with G = finger(inv(distortion)), per item."""
    out, scalar = _call(L.lib().hs_hank2016_stress, eos, den, pressure, distortion, 9, 9, device)
    return out[0] if scalar else out
