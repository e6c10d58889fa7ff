"""Multi-GPU driver, one process per GPU (torch.distributed): slab domain decomposition of one
large 1-D grid with a one-cell halo, and partitioning of ensembles of independent problems.

The reference has no distributed layer (its only parallelism is `Threads.@threads` over cells,
main.jl:206,221); this is the B200 replacement of those two loops for grids that span GPUs.
torch is plumbing only: device memory, the CUDA stream, and NCCL through torch.distributed.
All arithmetic runs in libhyperelastic_b200.so through the device-pointer layer of the C ABI.

Per step and per rank:
    hsd_step                         fused kernel (reads lambda_max slot n%3, writes slot (n+1)%3)
    halo exchange                    first / last owned cell (nvar + HS_NAUX doubles) to the neighbours
    all_reduce(MAX) of slot (n+1)%3  -> dt of the next step is bit-identical on all ranks and
                                        identical to the single-GPU run (max is exact)
Both collectives are enqueued on the same stream as the kernels: no host synchronisation.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L

__all__ = ["slab_bounds", "CudaKernels", "SlabSolver", "EnsembleSolver"]


def slab_bounds(n_global: int, world: int, rank: int):
    """Owned global cells [a, b) of `rank`, and the local array range [lo, hi) = owned cells plus
    one halo cell on each side that has a neighbour.  The physical boundary cells (global 0 and
    n-1, frozen by main.jl:219-220) are owned by the first / last rank and are the first / last
    local cell there."""
    def cut(k):   # global index of the first cell of rank k
        if k <= 0 or k >= world:
            return 0 if k <= 0 else n_global
        c = n_global * k // world
        # Odd interior cuts make every local array (owned cells + halo cells) even-sized when n_global is even: the first and
        # last slab own an odd number of cells and carry one halo cell, the others own an even number and carry two.  An even
        # row pitch is what the tensor-map tile copies of the single-phase step need (odd pitches take the row copies, ~6 % slower).
        if c % 2 == 0 and n_global // world >= 64:
            c -= 1
        return c
    a, b = cut(rank), cut(rank + 1)
    lo = a - (1 if rank > 0 else 0)
    hi = b + (1 if rank < world - 1 else 0)
    return a, b, lo, hi


class CudaKernels:
    """The product kernels behind the device-pointer ABI (hsd_*), on torch's current stream."""

    def __init__(self, eos, model, device):
        self._lib = L.lib()             # (also fills L.NAUX with the layout the library was built with)
        self.model = model
        self.nvar = L.NVAR[model]
        self.naux = L.NAUX[model]
        self.device = torch.device(device)
        self._eos = L.eos_array(eos, model)
        if self._lib.hs_device_count() <= 0:
            raise L.HyperelasticError(L.HS_ERR_CUDA, "no CUDA device visible: this library has no CPU fallback")

    def problem(self, ncells, nprob=1):
        p = L.HsdProblem()
        L.check(self._lib.hsd_problem_init(C.byref(p), self.model, self._eos, L.NPHASE[self.model], ncells, nprob))
        return p

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.device)

    def zeros(self, *shape):
        return torch.zeros(*shape, dtype=torch.float64, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def aos_to_soa(self, prob, aos, soa):
        L.check(self._lib.hsd_aos_to_soa(C.byref(prob), aos.data_ptr(), soa.data_ptr(), self._stream()))

    def soa_to_aos(self, prob, soa, aos):
        L.check(self._lib.hsd_soa_to_aos(C.byref(prob), soa.data_ptr(), aos.data_ptr(), self._stream()))

    def wave_bounds(self, prob, Q, aux, scal, slot, accumulate=False):
        fn = self._lib.hsd_wave_bounds_acc if accumulate else self._lib.hsd_wave_bounds
        L.check(fn(C.byref(prob), Q.data_ptr(), aux.data_ptr(), scal.data_ptr(), slot, self._stream()))

    def window(self, prob, ncells):
        """descriptor of a window of `ncells` cells of the grid `prob` describes (same row pitch): see WINDOWS in the header"""
        w = L.HsdProblem()
        C.memmove(C.byref(w), C.byref(prob), C.sizeof(L.HsdProblem))
        w.ncells, w.nprob = int(ncells), 1
        return w

    def step(self, prob, flux, cfl, dx, t_end, n, Qin, aux_in, Qout, aux_out, scal, ghost_mask, dt_hist=None, hist_k=0, hist_cap=0):
        L.check(self._lib.hsd_step(C.byref(prob), flux, cfl, dx, t_end, n, Qin.data_ptr(), aux_in.data_ptr(),
                                   Qout.data_ptr(), aux_out.data_ptr(), scal.data_ptr(),
                                   dt_hist.data_ptr() if dt_hist is not None else None, hist_k, hist_cap, ghost_mask, self._stream()))

    def halo(self, prob, Q, aux, left, right, mask, unpack):
        L.check(self._lib.hsd_halo(C.byref(prob), Q.data_ptr(), aux.data_ptr(), left.data_ptr(), right.data_ptr(),
                                   mask, int(unpack), self._stream()))

    def mailbox_doubles(self):
        return int(self._lib.hsd_mailbox_doubles())

    def exchange_p2p(self, prob, Q, aux, lam_slot, peer_ptrs, rank, world, seq, scal):
        arr = (C.c_void_p * world)(*[C.c_void_p(int(x)) for x in peer_ptrs])
        L.check(self._lib.hsd_exchange_p2p(C.byref(prob), Q.data_ptr(), aux.data_ptr(), lam_slot.data_ptr(), arr, rank, world, seq,
                                           scal.data_ptr(), self._stream()))

    def launches(self):
        return int(self._lib.hs_kernel_launch_count())


def scal_size(nprob):
    return L.HS_SCAL_SLOTS * nprob + 8


class _Base:
    def _views(self):
        np_ = self.nprob
        s = self.scal
        self._lam = [s[k * np_:(k + 1) * np_] for k in range(3)]          # lambda_max slots
        self._t = [s[3 * np_ + k * np_:3 * np_ + (k + 1) * np_] for k in range(3)]
        self._steps = s[6 * np_:7 * np_].view(torch.int64)
        self._status = s[L.HS_SCAL_SLOTS * np_:L.HS_SCAL_SLOTS * np_ + 1].view(torch.int32)

    @property
    def t(self):
        return self._t[self.n % 3].cpu().numpy().copy()

    @property
    def steps(self):
        return self._steps.cpu().numpy().copy()

    @property
    def lambda_max(self):
        return self._lam[self.n % 3].cpu().numpy().copy()

    def check_status(self):
        st = int(self._status.cpu()[0])
        if st & 2:
            raise L.HyperelasticError(L.HS_ERR_EXCHANGE, "peer-memory exchange timed out: another rank stopped stepping")
        if st & 4:
            raise L.HyperelasticError(L.HS_ERR_CUDA, "tile copy (TMA) did not complete: internal error of the single-phase step kernel")
        if st != 0:
            raise L.DomainError(L.HS_ERR_DOMAIN, "unphysical state (negative det / NaN): Julia would throw DomainError")


class SlabSolver(_Base):
    """One global grid of `n_global` cells split into contiguous slabs, one per rank."""

    def __init__(self, kernels, n_global, group=None):
        self.k = kernels
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_global = int(n_global)
        self.nvar = kernels.nvar
        self.nprob = 1
        self.a, self.b, self.lo_g, self.hi_g = slab_bounds(self.n_global, self.world, self.rank)
        self.nloc = self.hi_g - self.lo_g
        if self.nloc < 3:
            raise ValueError("slab too small: every rank needs at least 3 local cells")
        self.ghost_mask = (1 if self.rank > 0 else 0) | (2 if self.rank < self.world - 1 else 0)
        self.prob = kernels.problem(self.nloc, 1)
        self.Q = [kernels.empty(self.nvar, self.nloc) for _ in range(2)]
        self.aux = [kernels.empty(kernels.naux, self.nloc) for _ in range(2)]   # cached per-cell rows (wave bounds, ...)
        self.scal = kernels.zeros(scal_size(1))
        self._views()
        self.n = 0
        w = self.nvar + kernels.naux
        self._send = [kernels.empty(w), kernels.empty(w)]   # to left, to right
        self._recv = [kernels.empty(w), kernels.empty(w)]   # from left, from right
        self.exchange = "nccl"      # halo send/recv + all-reduce(max) through torch.distributed
        self._xseq = 0
        self._setup_p2p()

    def _setup_p2p(self):
        """One-kernel exchange over NVLink peer memory (hsd_exchange_p2p): mailboxes in symmetric memory.
        Falls back to the NCCL path when symmetric memory is unavailable (gloo, > 8 ranks, HS_EXCHANGE=nccl)."""
        import os
        if self.world == 1 or self.world > 8 or os.environ.get("HS_EXCHANGE", "p2p") != "p2p" or not hasattr(self.k, "exchange_p2p"):
            return
        if self.k.device.type != "cuda" or dist.get_backend(self.group) != "nccl":
            return
        try:
            import torch.distributed._symmetric_memory as symm
            self._mbox = symm.empty(self.k.mailbox_doubles(), dtype=torch.float64, device=self.k.device)
            self._mbox.zero_()
            hdl = symm.rendezvous(self._mbox, self.group if self.group is not None else dist.group.WORLD)
            self._peer_ptrs = [int(x) for x in hdl.buffer_ptrs]
            torch.cuda.synchronize(self.k.device)
            dist.barrier(group=self.group)       # every mailbox is zeroed before anybody posts
            ok = torch.ones(1, device=self.k.device)
        except Exception as e:                    # noqa: BLE001 -- any failure means "use NCCL"
            self._p2p_error = repr(e)
            ok = torch.zeros(1, device=self.k.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)   # all ranks must agree on the path
        if float(ok.item()) == 1.0:
            self.exchange = "p2p-kernel"

    # -- state ----------------------------------------------------------------------------------
    def set_local(self, Q_local):
        """Q_local: (nloc, nvar) Julia-layout rows of global cells [lo_g, hi_g) (halo included)."""
        aos = torch.as_tensor(np.ascontiguousarray(Q_local, dtype=np.float64)).to(self.k.device)
        self.set_local_device(aos)

    def set_local_device(self, aos, check=True):
        """aos: device tensor (nloc, nvar)."""
        assert tuple(aos.shape) == (self.nloc, self.nvar)
        self.scal.zero_()
        self.n = 0
        self.k.aos_to_soa(self.prob, aos, self.Q[0])
        self.k.wave_bounds(self.prob, self.Q[0], self.aux[0], self.scal, 0)
        self._allreduce_lambda(0)
        if check:
            self.check_status()

    def init_from_soa(self):
        """Q[0] was filled in place (structure of arrays): reset the clock and run the CFL sweep."""
        self.scal.zero_()
        self.n = 0
        self.k.wave_bounds(self.prob, self.Q[0], self.aux[0], self.scal, 0)
        self._allreduce_lambda(0)
        self.check_status()

    def local_aos_host(self):
        """(nloc, nvar) host copy of the local slab, halo cells included."""
        aos = self.k.empty(self.nloc, self.nvar)
        self.k.soa_to_aos(self.prob, self.Q[self.n & 1], aos)
        return aos.cpu().numpy()

    def step_host_serial(self, host_in, host_out, flux=L.HLL, cfl=0.6, dx=None):
        """One step on host slabs (pinned torch tensors (nloc, nvar)): H2D, CFL sweep (+ allreduce),
        fused step, halo, D2H one after the other."""
        if not hasattr(self, "_stage"):
            self._stage = self.k.empty(self.nloc, self.nvar)
        self._stage.copy_(host_in, non_blocking=True)
        self.set_local_device(self._stage, check=False)
        self.step(flux, cfl, dx)
        self.k.soa_to_aos(self.prob, self.Q[self.n & 1], self._stage)
        host_out.copy_(self._stage, non_blocking=True)
        torch.cuda.current_stream(self.k.device).synchronize()
        self._have_hint = True

    def step_host(self, host_in, host_out, flux=L.HLL, cfl=0.6, dx=None, chunk=1 << 19):
        """One step on host slabs, chunk-pipelined -- the multi-GPU counterpart of the C ABI's hs_step_host (see the header):
        H2D of chunk i+1 || transpose + CFL sweep + fused step of chunk i (speculative dt from the GLOBAL max(lambda) the previous
        call's exchange left in the scalar slot) || D2H of chunk i-1 on three streams; then the exchange (halo cells of the new
        state, global max(lambda) of the new state) and one all-reduce(max) of the sweep's max(lambda), compared with the hint bit
        for bit on every rank.  A refuted hint (first call, edited state) redoes the step on the device.  Results are bit-identical
        to step_host_serial."""
        import os
        dx = 1.0 / self.n_global if dx is None else dx
        N, nvar, dev = self.nloc, self.nvar, self.k.device
        if os.environ.get("HS_HOST_PIPELINE", "1") == "0" or N % 2 or N < 4096 or not host_in.is_pinned() or not host_out.is_pinned():
            return self.step_host_serial(host_in, host_out, flux, cfl, dx)
        if not hasattr(self, "_pipe"):
            self._stage = self.k.empty(N, nvar)
            self._pipe = dict(out=self.k.empty(N, nvar), sweep=self.k.zeros(scal_size(1)), s_in=torch.cuda.Stream(dev), s_out=torch.cuda.Stream(dev),
                              hint=self.k.zeros(1))
        P = self._pipe
        main = torch.cuda.current_stream(dev)
        # chunk boundaries: short first chunks, then equal ones; all even (hs_step_host uses the same rule)
        bnd, lo, sz = [], 0, max(chunk // 32, 2)
        while sz < chunk and N - lo > 2 * chunk + 4 * sz:
            bnd.append(lo); lo += sz & ~1; sz *= 2
        bnd += list(range(lo, N, chunk))
        if len(bnd) > 1 and N - bnd[-1] < chunk // 4:
            bnd.pop()
        bnd.append(N)
        K = len(bnd) - 1
        spec = bool(getattr(self, "_have_hint", False))
        self._have_hint = False
        if spec:
            P["hint"].copy_(self._lam[self.n % 3])          # global max(lambda) of the state the previous call returned
        self.scal.zero_(); P["sweep"].zero_()
        self.n = 0
        if spec:
            self._lam[0].copy_(P["hint"])
        start = torch.cuda.Event(); start.record(main)
        P["s_in"].wait_event(start); P["s_out"].wait_event(start)
        Q0, Q1, A0, A1 = self.Q[0], self.Q[1], self.aux[0], self.aux[1]
        for i in range(K):
            b0, b1 = bnd[i], bnd[i + 1]
            with torch.cuda.stream(P["s_in"]):
                self._stage[b0:b1].copy_(host_in[b0:b1], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(P["s_in"])
            main.wait_event(ev)
            w = self.k.window(self.prob, b1 - b0)
            self.k.aos_to_soa(w, self._stage[b0:b1], Q0[:, b0:])
            self.k.wave_bounds(w, Q0[:, b0:], A0[:, b0:], P["sweep"], 0, accumulate=True)
            if spec:
                w0 = b0 - 2 if i else 0
                ghost = ((1 if i else (self.ghost_mask & 1)) | (2 if i < K - 1 else (self.ghost_mask & 2)))
                ws = self.k.window(self.prob, b1 - w0)
                self.k.step(ws, flux, cfl, dx, 1.0e300, 0, Q0[:, w0:], A0[:, w0:], Q1[:, w0:], A1[:, w0:], self.scal, ghost)
                u0, u1 = (b0 - 1 if i else 0), (b1 - 1 if i < K - 1 else N)
                wo = self.k.window(self.prob, u1 - u0)
                self.k.soa_to_aos(wo, Q1[:, u0:], P["out"][u0:u1])
                ev2 = torch.cuda.Event(); ev2.record(main)
                P["s_out"].wait_event(ev2)
                with torch.cuda.stream(P["s_out"]):
                    host_out[u0:u1].copy_(P["out"][u0:u1], non_blocking=True)
        lam_true = P["sweep"][0:1]
        if self.world > 1:
            dist.all_reduce(lam_true, op=dist.ReduceOp.MAX, group=self.group)
        ok = spec and bool(torch.equal(lam_true.view(torch.int64), P["hint"].view(torch.int64)))   # (host sync; identical on all ranks)
        self._status |= P["sweep"][L.HS_SCAL_SLOTS:L.HS_SCAL_SLOTS + 1].view(torch.int32)
        if not ok:
            P["s_out"].synchronize()
            st = self._status.clone()
            self.scal.zero_(); self._status.copy_(st)
            self._lam[0].copy_(lam_true)
            self.n = 0
            self.step(flux, cfl, dx)                       # fused step + exchange from the intact input
            for i in range(K):
                b0, b1 = bnd[i], bnd[i + 1]
                wo = self.k.window(self.prob, b1 - b0)
                self.k.soa_to_aos(wo, Q1[:, b0:], P["out"][b0:b1])
                ev2 = torch.cuda.Event(); ev2.record(main)
                P["s_out"].wait_event(ev2)
                with torch.cuda.stream(P["s_out"]):
                    host_out[b0:b1].copy_(P["out"][b0:b1], non_blocking=True)
        else:
            # every window counted a step; the grid took one.  Then what step() does after the kernel: halo cells + global max(lambda)
            self._steps.fill_(1)
            if self.exchange == "p2p-kernel":
                self._xseq += 1
                self.k.exchange_p2p(self.prob, Q1, A1, self._lam[1], self._peer_ptrs, self.rank, self.world, self._xseq, self.scal)
            else:
                self._halo_exchange(1)
                self._allreduce_lambda(1)
            self.n = 1
            if self.ghost_mask:   # the halo cells of the new state arrived with the exchange: send them after the chunks
                for u0, u1 in ((0, 2), (N - 2, N)):
                    wo = self.k.window(self.prob, 2)
                    self.k.soa_to_aos(wo, Q1[:, u0:], P["out"][u0:u1])
                ev2 = torch.cuda.Event(); ev2.record(main)
                P["s_out"].wait_event(ev2)
                with torch.cuda.stream(P["s_out"]):
                    host_out[0:2].copy_(P["out"][0:2], non_blocking=True)
                    host_out[N - 2:N].copy_(P["out"][N - 2:N], non_blocking=True)
        main.synchronize(); P["s_out"].synchronize()
        self._have_hint = True
        self.pipelined_calls = getattr(self, "pipelined_calls", 0) + 1
        self.speculation_hits = getattr(self, "speculation_hits", 0) + (1 if ok else 0)

    def set_from_global(self, Q_global):
        self.set_local(np.asarray(Q_global)[self.lo_g:self.hi_g])

    def owned(self):
        """(a, b, Q_owned (b-a, nvar)) of the cells this rank owns."""
        aos = self.k.empty(self.nloc, self.nvar)
        self.k.soa_to_aos(self.prob, self.Q[self.n & 1], aos)
        off = self.a - self.lo_g
        return self.a, self.b, aos[off:off + (self.b - self.a)].cpu().numpy()

    def gather(self):
        """Global (n_global, nvar) array on every rank (test / IO helper; not on the hot path)."""
        a, b, mine = self.owned()
        if self.world == 1:
            return mine
        parts = [None] * self.world
        dist.all_gather_object(parts, (a, b, mine), group=self.group)
        out = np.empty((self.n_global, self.nvar))
        for (pa, pb, q) in parts:
            out[pa:pb] = q
        return out

    # -- communication ----------------------------------------------------------------------------
    def _allreduce_lambda(self, slot):
        if self.world > 1:
            dist.all_reduce(self._lam[slot], op=dist.ReduceOp.MAX, group=self.group)

    def _halo_exchange(self, buf):
        if self.world == 1:
            return
        Q, aux = self.Q[buf], self.aux[buf]
        self.k.halo(self.prob, Q, aux, self._send[0], self._send[1], self.ghost_mask, False)
        ops = []
        if self.rank > 0:
            ops += [dist.P2POp(dist.isend, self._send[0], self.rank - 1, self.group), dist.P2POp(dist.irecv, self._recv[0], self.rank - 1, self.group)]
        if self.rank < self.world - 1:
            ops += [dist.P2POp(dist.isend, self._send[1], self.rank + 1, self.group), dist.P2POp(dist.irecv, self._recv[1], self.rank + 1, self.group)]
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        self.k.halo(self.prob, Q, aux, self._recv[0], self._recv[1], self.ghost_mask, True)

    # -- the loop -------------------------------------------------------------------------------
    def step(self, flux=L.HLL, cfl=0.6, dx=None, t_end=1.0e300, kernel_events=None):
        dx = 1.0 / self.n_global if dx is None else dx
        a, b = self.n & 1, (self.n & 1) ^ 1
        if kernel_events:
            kernel_events[0].record()
        self.k.step(self.prob, flux, cfl, dx, t_end, self.n, self.Q[a], self.aux[a], self.Q[b], self.aux[b], self.scal, self.ghost_mask)
        if kernel_events:
            kernel_events[1].record()
        if self.exchange == "p2p-kernel":
            self._xseq += 1
            self.k.exchange_p2p(self.prob, self.Q[b], self.aux[b], self._lam[(self.n + 1) % 3], self._peer_ptrs, self.rank, self.world,
                                self._xseq, self.scal)
        else:
            self._halo_exchange(b)
            self._allreduce_lambda((self.n + 1) % 3)
        self.n += 1

    def advance(self, t_end, flux=L.HLL, cfl=0.6, dx=None, max_steps=1 << 30, check_every=32):
        done = 0
        while done < max_steps:
            if not (self.t[0] < t_end):
                break
            m = min(check_every, max_steps - done)
            for _ in range(m):
                self.step(flux, cfl, dx, t_end)
            done += m
        self.check_status()
        return done


class EnsembleSolver(_Base):
    """`nprob_global` independent problems of `ncells` cells partitioned over the ranks; every
    problem keeps its own dt / t.  No communication while stepping."""

    def __init__(self, kernels, ncells, nprob_global, group=None):
        self.k = kernels
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.p0 = nprob_global * self.rank // self.world
        self.p1 = nprob_global * (self.rank + 1) // self.world
        self.nprob = self.p1 - self.p0
        self.ncells, self.nvar = int(ncells), kernels.nvar
        self.prob = kernels.problem(self.ncells, self.nprob)
        tot = self.ncells * self.nprob
        self.Q = [kernels.empty(self.nvar, tot) for _ in range(2)]
        self.aux = [kernels.empty(kernels.naux, tot) for _ in range(2)]
        self.scal = kernels.zeros(scal_size(self.nprob))
        self._views()
        self.n = 0

    def set_local(self, Q_local):
        """(nprob_local, ncells, nvar) Julia layout."""
        aos = torch.as_tensor(np.ascontiguousarray(Q_local, dtype=np.float64)).to(self.k.device).reshape(-1, self.nvar)
        assert aos.shape[0] == self.ncells * self.nprob
        self.set_local_device(aos)

    def set_local_device(self, aos):
        self.scal.zero_()
        self.n = 0
        self.k.aos_to_soa(self.prob, aos, self.Q[0])
        self.k.wave_bounds(self.prob, self.Q[0], self.aux[0], self.scal, 0)
        self.check_status()

    def init_from_soa(self):
        self.scal.zero_()
        self.n = 0
        self.k.wave_bounds(self.prob, self.Q[0], self.aux[0], self.scal, 0)
        self.check_status()

    def local(self):
        aos = self.k.empty(self.ncells * self.nprob, self.nvar)
        self.k.soa_to_aos(self.prob, self.Q[self.n & 1], aos)
        return aos.cpu().numpy().reshape(self.nprob, self.ncells, self.nvar)

    def local_aos_host(self):
        return self.local().reshape(-1, self.nvar)

    def step_host(self, host_in, host_out, flux=L.HLL, cfl=0.6, dx=None):
        if not hasattr(self, "_stage"):
            self._stage = self.k.empty(self.ncells * self.nprob, self.nvar)
        self._stage.copy_(host_in, non_blocking=True)
        self.scal.zero_()
        self.n = 0
        self.k.aos_to_soa(self.prob, self._stage, self.Q[0])
        self.k.wave_bounds(self.prob, self.Q[0], self.aux[0], self.scal, 0)
        self.step(flux, cfl, dx)
        self.k.soa_to_aos(self.prob, self.Q[self.n & 1], self._stage)
        host_out.copy_(self._stage, non_blocking=True)
        torch.cuda.current_stream(self.k.device).synchronize()

    def step(self, flux=L.HLL, cfl=0.6, dx=None, t_end=1.0e300, kernel_events=None):
        dx = 1.0 / self.ncells if dx is None else dx
        a, b = self.n & 1, (self.n & 1) ^ 1
        if kernel_events:
            kernel_events[0].record()
        self.k.step(self.prob, flux, cfl, dx, t_end, self.n, self.Q[a], self.aux[a], self.Q[b], self.aux[b], self.scal, 0)
        if kernel_events:
            kernel_events[1].record()
        self.n += 1

    def advance(self, t_end, flux=L.HLL, cfl=0.6, dx=None, max_steps=1 << 30, check_every=32):
        done = 0
        while done < max_steps:
            if not (self.t < t_end).any():
                break
            m = min(check_every, max_steps - done)
            for _ in range(m):
                self.step(flux, cfl, dx, t_end)
            done += m
        self.check_status()
        return done
