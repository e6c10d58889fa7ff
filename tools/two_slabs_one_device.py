"""Development aid / round-2 check: the slab decomposition with ghost cells on ONE device.

`world` SlabSolver objects with faked ranks share a device; after every fused step their halo cells are exchanged with
hsd_halo pack / unpack + device copies and max(lambda) with an element-wise maximum -- the NCCL exchange without NCCL.  The
gathered result must be bit-identical to the single-domain run (only an exact max crosses slabs).  With the odd interior
cuts of slab_bounds every slab of an even-sized grid takes the tensor-map tile copies, so this exercises k_step_sp<TM2D>
with a ghost mask on a 1-GPU box (the multi-GPU tests need 2 GPUs and are skipped there).

  python tools/two_slabs_one_device.py            # on a GPU box: product kernels
  python tools/two_slabs_one_device.py --cpu      # here: the oracle-backed kernel double (checks this script's own logic)
"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import numpy as np
import torch
from hyperelasticsolver_b200 import _lib as L
from hyperelasticsolver_b200.slab import SlabSolver, slab_bounds, scal_size


def fake_rank_slab(kernels, n_global, world, rank):
    """A SlabSolver for (rank, world) without torch.distributed (same fields as SlabSolver.__init__)."""
    s = object.__new__(SlabSolver)
    s.k, s.group, s.world, s.rank, s.n_global, s.nvar, s.nprob = kernels, None, world, rank, int(n_global), kernels.nvar, 1
    s.a, s.b, s.lo_g, s.hi_g = slab_bounds(s.n_global, world, rank)
    s.nloc = s.hi_g - s.lo_g
    s.ghost_mask = (1 if rank > 0 else 0) | (2 if rank < world - 1 else 0)
    s.prob = kernels.problem(s.nloc, 1)
    s.Q = [kernels.empty(s.nvar, s.nloc) for _ in range(2)]
    s.aux = [kernels.empty(kernels.naux, s.nloc) for _ in range(2)]
    s.scal = kernels.zeros(scal_size(1))
    s._views()
    s.n = 0
    w = s.nvar + kernels.naux
    s._send = [kernels.empty(w), kernels.empty(w)]
    s._recv = [kernels.empty(w), kernels.empty(w)]
    s.exchange = "local"
    return s


def run_slabs(kernels, Q0, world, nsteps, flux, cfl=0.6):
    n = Q0.shape[0]
    dx = 1.0 / n
    slabs = [fake_rank_slab(kernels, n, world, r) for r in range(world)]

    def share_lambda(slot):
        m = slabs[0]._lam[slot].clone()
        for s in slabs[1:]:
            m = torch.maximum(m, s._lam[slot])
        for s in slabs:
            s._lam[slot].copy_(m)

    for s in slabs:
        aos = torch.as_tensor(np.ascontiguousarray(Q0[s.lo_g:s.hi_g])).to(kernels.device)
        s.k.aos_to_soa(s.prob, aos, s.Q[0])
        s.k.wave_bounds(s.prob, s.Q[0], s.aux[0], s.scal, 0)
    share_lambda(0)
    for _ in range(nsteps):
        for s in slabs:
            a, b = s.n & 1, (s.n & 1) ^ 1
            s.k.step(s.prob, flux, cfl, dx, 1e300, s.n, s.Q[a], s.aux[a], s.Q[b], s.aux[b], s.scal, s.ghost_mask)
            s.k.halo(s.prob, s.Q[b], s.aux[b], s._send[0], s._send[1], s.ghost_mask, False)
        for r, s in enumerate(slabs):
            if r > 0:
                s._recv[0].copy_(slabs[r - 1]._send[1])
            if r < world - 1:
                s._recv[1].copy_(slabs[r + 1]._send[0])
        for s in slabs:
            b = (s.n & 1) ^ 1
            s.k.halo(s.prob, s.Q[b], s.aux[b], s._recv[0], s._recv[1], s.ghost_mask, True)
        share_lambda((slabs[0].n + 1) % 3)
        for s in slabs:
            s.n += 1
    out = np.empty_like(Q0)
    for s in slabs:
        s.check_status()
        a, b, mine = s.owned()
        out[a:b] = mine
    return out, slabs[0].t[0], [s.nloc for s in slabs]


def main():
    cpu = "--cpu" in sys.argv
    import hyperelasticsolver_b200 as hs
    from hyperelasticsolver_b200.testcases import mph_primitive_states, riemann_grid, sp_primitive_states
    ok = True
    for model, nx, world, nsteps in ((0, 20000, 2, 12), (0, 20001, 2, 12), (0, 40000, 4, 12), (1, 3000, 3, 6)) if not cpu else ((0, 132, 2, 5), (0, 400, 4, 5), (1, 135, 2, 3)):
        if cpu:
            import oracle as O
            from oracle_kernels import OracleKernels
            eos = [O.barton2009()] * (2 if model else 1)
            Pl, Pr = mph_primitive_states(6) if model else sp_primitive_states(1)
            Qlr, _ = O.prim2cons(eos, model, np.stack([Pl, Pr]))
            Q0 = riemann_grid(Qlr[0], Qlr[1], nx)
            kern = OracleKernels(eos, model)
            ref = O.run(eos, model, O.HLL, Q0, 0.6, 1.0 / nx, 1e9, nsteps)["Q"]
        else:
            from hyperelasticsolver_b200.slab import CudaKernels
            if model:
                eos = (hs.Barton2009(), hs.Barton2009()); Ql, Qr = hs.initial_states(eos, 6)
            else:
                eos = hs.Barton2009(); Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
            Q0 = hs.initial_condition(Ql, Qr, nx)
            kern = CudaKernels(eos, hs.MPH30 if model else hs.SP13, "cuda:0")
            with hs.Solver(eos, nx, model=hs.MPH30 if model else hs.SP13) as s1:
                s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps); ref = s1.download()
        Q, t, nlocs = run_slabs(kern, Q0, world, nsteps, L.HLL)
        same = np.array_equal(Q, ref)
        ok &= same
        print(f"model={'mph30' if model else 'sp13'} nx={nx} slabs={world} local sizes={nlocs} steps={nsteps}: "
              f"{'bit-identical to the single-domain run' if same else 'DIFFERS: max |d| = %.3e' % np.abs(Q - ref).max()}", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
