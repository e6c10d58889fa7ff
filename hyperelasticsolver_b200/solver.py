"""The time loop of main.jl:202-241 behind the C ABI: a device-resident solver object plus the
host-side helpers of main.jl that do not touch physics (`initial_condition`, main.jl:99-106).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .testcases import riemann_grid

__all__ = ["Solver", "Solver2D", "initial_condition", "update_cell", "register_host", "unregister_host"]

_FLUX = {"hll": L.HLL, "lxf": L.LXF, L.HLL: L.HLL, L.LXF: L.LXF}


def initial_condition(Ql, Qr, nx):
    """main.jl:99-106 -> (nx, nvar), byte-identical to Julia's (nvar, nx) column-major array."""
    return riemann_grid(Ql, Qr, nx)


class Solver:
    """Owns the device copy of Q0 (structure-of-arrays, double-buffered) for `nprob` independent
    problems of `ncells` cells.  One instance replaces the body of `while t < T` (main.jl:202-227).
    """

    def __init__(self, eos, ncells, nprob=1, model=L.MPH30, device=0, devices=None):
        """devices=[0, 1, ...]: spread the work over several GPUs of this process (hs_create_multi): one grid
        (nprob == 1) is slab-decomposed, an ensemble (nprob > 1) is shared out by problems; results are
        bit-identical to the single-device solver."""
        self.model, self.nvar = model, L.NVAR[model]
        self.ncells, self.nprob, self.device = int(ncells), int(nprob), int(device)
        self._ctx = C.c_void_p()
        self.t = np.zeros(self.nprob)
        self.steps = np.zeros(self.nprob, dtype=np.int64)
        self._eos = L.eos_array(eos, model)
        if devices is not None and len(devices) > 1:
            devs = (C.c_int * len(devices))(*[int(d) for d in devices])
            L.check(L.lib().hs_create_multi(C.byref(self._ctx), model, self._eos, L.NPHASE[model], self.ncells, self.nprob, devs,
                                            len(devices)))
        else:
            if devices:
                self.device = int(devices[0])
            L.check(L.lib().hs_create(C.byref(self._ctx), model, self._eos, L.NPHASE[model], self.ncells, self.nprob, self.device))

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            L.lib().hs_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state --------------------------------------------------------------------------------
    def _shape(self):
        return (self.ncells, self.nvar) if self.nprob == 1 else (self.nprob, self.ncells, self.nvar)

    def upload(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.size != self.nprob * self.ncells * self.nvar:
            raise ValueError(f"expected {self._shape()} values, got {Q.shape}")
        L.check(L.lib().hs_upload(self._ctx, Q.ctypes.data))
        self.t = np.zeros(self.nprob)
        self.steps = np.zeros(self.nprob, dtype=np.int64)
        return self

    def download(self, out=None):
        Q = np.empty(self._shape()) if out is None else out
        L.check(L.lib().hs_download(self._ctx, Q.ctypes.data))
        return Q

    def set_time(self, t, step=0):
        """restart: main.jl:185-186"""
        self.t[:] = t
        self.steps[:] = step
        L.check(L.lib().hs_set_time(self._ctx, float(t), int(step)))

    # -- the loop -----------------------------------------------------------------------------
    def wave_speeds(self, full=False):
        """CFL sweep main.jl:204-212 -> lambda_max per problem (and get_eigvals of every cell)."""
        lam = np.empty(self.nprob)
        eig = np.empty(self._shape()[:-1] + (6 * L.NPHASE[self.model],)) if full else None
        L.check(L.lib().hs_wave_speeds(self._ctx, eig.ctypes.data if full else None, lam.ctypes.data))
        return (lam, eig) if full else lam

    def step(self, flux="hll", cfl=0.6, dx=None):
        """One iteration of main.jl:204-227; returns dt per problem."""
        dx = 1.0 / self.ncells if dx is None else dx
        dt = np.empty(self.nprob)
        L.check(L.lib().hs_step(self._ctx, _FLUX[flux], float(cfl), float(dx), dt.ctypes.data))
        self.t += dt          # main.jl:214
        self.steps += 1       # main.jl:215
        return dt

    def advance(self, t_end, flux="hll", cfl=0.6, dx=None, max_steps=1 << 30, record_dt=False):
        """`while t < T` (main.jl:202): device-resident until every problem has t >= t_end (no
        clipping of the last step) or max_steps more steps were taken.  Updates self.t / self.steps;
        returns the (nprob, max_steps) dt history if record_dt."""
        dx = 1.0 / self.ncells if dx is None else dx
        if record_dt and max_steps > (1 << 24):
            raise ValueError("record_dt needs a finite max_steps")
        hist = np.zeros((self.nprob, max_steps)) if record_dt else None
        L.check(L.lib().hs_advance(self._ctx, _FLUX[flux], float(cfl), float(dx), float(t_end), int(max_steps),
                                   self.t.ctypes.data, self.steps.ctypes.data, hist.ctypes.data if record_dt else None))
        return hist

    def step_host(self, Qin, Qout=None, flux="hll", cfl=0.6, dx=None):
        """One step on host arrays (upload + step + download), the literal drop-in for one pass
        of main.jl:204-227 with Q0 in host memory."""
        dx = 1.0 / self.ncells if dx is None else dx
        Qin = np.ascontiguousarray(Qin, dtype=np.float64)
        Qout = np.empty_like(Qin) if Qout is None else Qout
        dt = np.empty(self.nprob)
        L.check(L.lib().hs_step_host(self._ctx, _FLUX[flux], float(cfl), float(dx), Qin.ctypes.data, Qout.ctypes.data, dt.ctypes.data))
        self.t = dt.copy()                                   # the context is left as upload + one step leave it
        self.steps = np.ones(self.nprob, dtype=np.int64)
        return Qout, dt

    def step_host_stats(self):
        """(calls of step_host that took the chunk-pipelined form, how many of them had their hinted max(lambda) confirmed)"""
        a, b = C.c_int64(0), C.c_int64(0)
        L.check(L.lib().hs_step_host_stats(self._ctx, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)


class Solver2D:
    """Dimension-split 2-D solver on an nx x ny grid (hs2d_*): Q^{n+1} = Y(dt) X(dt) Q^n with the 1-D step of
    main.jl:204-227 along rows and columns, dt = cfl min(dx / max lambda_x, dy / max lambda_y).  Arrays are (ny, nx, nvar)
    C-contiguous = Julia's Array{Float64,3}(nvar, nx, ny)."""

    def __init__(self, eos, nx, ny, model=L.MPH30, device=0):
        self.model, self.nvar, self.nx, self.ny = model, L.NVAR[model], int(nx), int(ny)
        self._ctx = C.c_void_p()
        self._eos = L.eos_array(eos, model)
        L.check(L.lib().hs2d_create(C.byref(self._ctx), model, self._eos, L.NPHASE[model], self.nx, self.ny, int(device)))
        self.t, self.steps = 0.0, 0

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            L.lib().hs2d_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def upload(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.shape != (self.ny, self.nx, self.nvar):
            raise ValueError(f"expected {(self.ny, self.nx, self.nvar)}, got {Q.shape}")
        L.check(L.lib().hs2d_upload(self._ctx, Q.ctypes.data))
        self.t, self.steps = 0.0, 0
        return self

    def download(self):
        Q = np.empty((self.ny, self.nx, self.nvar))
        L.check(L.lib().hs2d_download(self._ctx, Q.ctypes.data))
        return Q

    def step(self, flux="hll", cfl=0.6, dx=None, dy=None):
        dx = 1.0 / self.nx if dx is None else dx
        dy = 1.0 / self.ny if dy is None else dy
        dt = C.c_double(0.0)
        L.check(L.lib().hs2d_step(self._ctx, _FLUX[flux], float(cfl), float(dx), float(dy), C.byref(dt)))
        self.t += dt.value; self.steps += 1
        return dt.value

    def advance(self, t_end, flux="hll", cfl=0.6, dx=None, dy=None, max_steps=1 << 30):
        dx = 1.0 / self.nx if dx is None else dx
        dy = 1.0 / self.ny if dy is None else dy
        t, n = C.c_double(0.0), C.c_int64(0)
        L.check(L.lib().hs2d_advance(self._ctx, _FLUX[flux], float(cfl), float(dx), float(dy), float(t_end), int(max_steps), C.byref(t), C.byref(n)))
        self.t, self.steps = t.value, int(n.value)
        return self.steps


def register_host(arr):
    """Page-lock a numpy array (hs_host_register) so that upload / download / step_host copy at the link rate and the chunks of
    step_host overlap; returns the array.  Release with unregister_host."""
    L.check(L.lib().hs_host_register(arr.ctypes.data, arr.nbytes))
    return arr


def unregister_host(arr):
    L.check(L.lib().hs_host_unregister(arr.ctypes.data))


def update_cell(Q3, flux_num, eigvals_or_lambda, dtdx_or_eos, eos=None, device=0):
    """main.jl:30-60, both methods, on one 3-cell stencil `Q3` of shape (3, nvar):
      update_cell(Q3, lxf, lambda, eos)                  -> main.jl:30-41
      update_cell(Q3, hll, eigvals[3], dtdx, eos)        -> main.jl:43-60
    """
    Q3 = np.asarray(Q3, dtype=np.float64)
    Q_l, Q, Q_r = Q3[0], Q3[1], Q3[2]
    if eos is None:   # LxF method
        lam, eos = float(eigvals_or_lambda), dtdx_or_eos
        F_l, _, NF_l = flux_num(eos, Q_l, Q, lam, device=device)
        F_r, NF_r, _ = flux_num(eos, Q, Q_r, lam, device=device)
        return Q - 1.0 / lam * ((F_r - F_l) + (NF_r + NF_l))
    eig, dtdx = eigvals_or_lambda, float(dtdx_or_eos)
    F_l, _, NF_l = flux_num(eos, Q_l, Q, eig[0:2], device=device)
    F_r, NF_r, _ = flux_num(eos, Q, Q_r, eig[1:3], device=device)
    return Q - dtdx * ((F_r - F_l) + (NF_r + NF_l))
