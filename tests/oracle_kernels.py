"""Test double for hyperelasticsolver_b200.slab.CudaKernels: the same interface executed on CPU
tensors with the oracle, so the host-side slab logic (partition, halo exchange, all-reduce of
lambda_max, ghost handling, scalar-slot rotation) can be exercised with gloo on a machine
without a GPU.  Test infrastructure only."""
import types

import numpy as np
import torch

import oracle as O


class OracleKernels:
    def __init__(self, eos_blocks, model):
        self.model = model
        self.nvar = O.NVAR[model]
        # same number of cached rows as the product layout (hsd_naux of the library as built), so halo buffers and
        # mailboxes have the product's width; the double keeps lo / hi in rows 0, 1 and carries the rest unused
        from hyperelasticsolver_b200 import _lib as L
        L.lib()
        self.naux = L.NAUX[model]
        assert self.naux >= 2
        self.neig = O.NEIG[model]
        self.eos = eos_blocks
        self.device = torch.device("cpu")
        self._launches = 0

    def problem(self, ncells, nprob=1):
        return types.SimpleNamespace(ncells=ncells, nprob=nprob)

    def empty(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64)

    zeros = empty

    def aos_to_soa(self, prob, aos, soa):
        soa.copy_(aos.T)

    def soa_to_aos(self, prob, soa, aos):
        aos.copy_(soa.T)

    def launches(self):
        return self._launches

    def halo(self, prob, Q, aux, left, right, mask, unpack):
        nv, last = self.nvar, Q.shape[1] - 1
        for side, buf in ((0, left), (1, right)):
            if not (mask & (1 << side)):
                continue
            if unpack:
                c = last if side else 0
                Q[:, c] = buf[:nv]; aux[:, c] = buf[nv:]
            else:
                c = last - 1 if side else 1
                buf[:nv] = Q[:, c]; buf[nv:] = aux[:, c]

    def _bounds(self, Qaos):
        eig, st = O.get_eigvals(self.eos, self.model, Qaos)
        assert st == 0
        return eig.min(axis=1), eig.max(axis=1), eig

    def wave_bounds(self, prob, Q, aux, scal, slot):
        assert prob.nprob == 1
        lo, hi = aux[0], aux[1]
        l, h, _ = self._bounds(Q.numpy().T.copy())
        lo.copy_(torch.from_numpy(l)); hi.copy_(torch.from_numpy(h))
        scal[slot] = float(np.maximum(np.abs(l), np.abs(h)).max())
        self._launches += 1

    def step(self, prob, flux, cfl, dx, t_end, n, Qin, aux_in, Qout, aux_out, scal, ghost_mask, **kw):
        """main.jl:212-227 on the local array; first / last cell frozen or ghost."""
        assert prob.nprob == 1
        lo_in, hi_in, lo_out, hi_out = aux_in[0], aux_in[1], aux_out[0], aux_out[1]
        cur, nxt, clr = n % 3, (n + 1) % 3, (n + 2) % 3
        t_cur = float(scal[3 + cur])
        lam = float(scal[cur])
        self._launches += 1
        if not (t_cur < t_end):
            Qout.copy_(Qin); lo_out.copy_(lo_in); hi_out.copy_(hi_in)
            scal[3 + nxt] = t_cur; scal[nxt] = lam; scal[clr] = 0.0
            return
        dt = cfl * dx / lam
        Q = Qin.numpy().T.copy()
        nc = Q.shape[0]
        # cached speeds of the cells: only min / max are read by hll (NumFluxes.jl:90-91)
        el = np.repeat(lo_in.numpy()[:, None], self.neig, 1); er = np.repeat(hi_in.numpy()[:, None], self.neig, 1)
        eig_l = np.where(np.arange(self.neig)[None, :] == 0, el, er)   # a vector whose min is lo and max is hi
        Ql, Qr = Q[:-1], Q[1:]
        if self.model == O.MPH30:
            if flux == O.HLL:
                cons, dm, dp, s, st = O.hll(self.eos, Ql, Qr, eig_l[:-1], eig_l[1:])
                Qn = Q[1:-1] - (dt / dx) * ((cons[1:] - cons[:-1]) + (dm[1:] + dp[:-1]))
            else:
                lamb = dx / dt
                cons, dm, dp, st = O.lxf(self.eos, Ql, Qr, lamb)
                Qn = Q[1:-1] - 1.0 / lamb * ((cons[1:] - cons[:-1]) + (dm[1:] + dp[:-1]))
        else:
            assert flux == O.HLL
            cons, st = O.sp_hll(self.eos, Ql, Qr, eig_l[:-1], eig_l[1:])
            Qn = Q[1:-1] - (dt / dx) * (cons[1:] - cons[:-1])
        assert st == 0
        Qnew = Q.copy(); Qnew[1:-1] = Qn
        l, h, _ = self._bounds(Qnew[1:-1])
        lo_new = lo_in.numpy().copy(); hi_new = hi_in.numpy().copy()
        lo_new[1:-1] = l; hi_new[1:-1] = h
        lamv = np.maximum(np.abs(lo_new), np.abs(hi_new))
        if ghost_mask & 1: lamv[0] = 0.0
        if ghost_mask & 2: lamv[-1] = 0.0
        Qout.copy_(torch.from_numpy(Qnew.T.copy())); lo_out.copy_(torch.from_numpy(lo_new)); hi_out.copy_(torch.from_numpy(hi_new))
        scal[nxt] = max(float(scal[nxt]), float(lamv.max()))
        scal[3 + nxt] = t_cur + dt
        st_i = scal[6:7].view(torch.int64); st_i += 1
        scal[clr] = 0.0
