"""CPU, world_size 2, gloo: the slab decomposition host logic (hyperelasticsolver_b200/slab.py)
driven by the oracle-backed kernel double must reproduce the single-domain oracle run exactly
(max-reduction is exact and no other arithmetic crosses ranks)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, model, flux, nx, nsteps, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle as O
        from oracle_kernels import OracleKernels
        from hyperelasticsolver_b200.slab import SlabSolver, slab_bounds
        from hyperelasticsolver_b200.testcases import mph_primitive_states, riemann_grid, sp_primitive_states
        if model == O.MPH30:
            eos = [O.barton2009(), O.barton2009()]
            Pl, Pr = mph_primitive_states(6)
        else:
            eos = [O.barton2009()]
            Pl, Pr = sp_primitive_states(1)
        Qlr, _ = O.prim2cons(eos, model, np.stack([Pl, Pr]))
        Q0 = riemann_grid(Qlr[0], Qlr[1], nx)
        sol = SlabSolver(OracleKernels(eos, model), nx)
        a, b, lo, hi = slab_bounds(nx, world, rank)
        assert (sol.a, sol.b) == (a, b) and sol.nloc == hi - lo
        assert sol.ghost_mask == ((1 if rank > 0 else 0) | (2 if rank < world - 1 else 0))
        sol.set_from_global(Q0)
        dts = []
        for _ in range(nsteps):
            lam = sol.lambda_max[0]
            dts.append(0.6 * (1.0 / nx) / lam)
            sol.step(flux, 0.6, 1.0 / nx)
        Q = sol.gather()
        if rank == 0:
            ref = O.run(eos, model, flux, Q0, 0.6, 1.0 / nx, 1e9, nsteps)
            q.put((np.array_equal(Q, ref["Q"]), float(np.abs(Q - ref["Q"]).max()), np.array_equal(np.array(dts), ref["dt"][0]),
                   float(sol.t[0]), float(ref["t"][0]), int(sol.steps[0])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,flux,nx", [(1, 1, 41), (0, 1, 41), (1, 0, 41), (0, 1, 132)])   # 132: shifted (odd) interior cut
def test_two_rank_slab_matches_single_domain(model, flux, nx):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nsteps = 6
    procs = [ctx.Process(target=_worker, args=(r, 2, port, model, flux, nx, nsteps, q)) for r in range(2)]
    for p in procs: p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    same, maxdiff, dt_same, t, t_ref, steps = res
    assert same, f"slab-decomposed run differs from the single-domain run by {maxdiff:.3e}"
    assert dt_same and t == t_ref and steps == nsteps


def test_slab_bounds_cover_domain():
    from hyperelasticsolver_b200.slab import slab_bounds
    for n, w in [(41, 2), (1000, 8), (1 << 20, 4), (17, 5), (1 << 25, 2), (1 << 27, 8), (1 << 28, 8), (1 << 28, 1), (130, 2)]:
        prev = 0
        for r in range(w):
            a, b, lo, hi = slab_bounds(n, w, r)
            assert a == prev and b > a
            assert lo == a - (r > 0) and hi == b + (r < w - 1)
            assert abs((b - a) - n / w) <= 2            # balanced to within a cell or two
            if n % 2 == 0 and n // w >= 64:
                assert (hi - lo) % 2 == 0                # even local arrays: the tensor-map tile copies apply on every rank
            prev = b
        assert prev == n
