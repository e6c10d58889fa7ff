"""GPU: the torch-plumbed device-pointer path (hyperelasticsolver_b200/slab.py) -- the one
bench.py and the multi-GPU driver use -- against the oracle and against the C-ABI context."""
import os
import socket
import sys

import numpy as np
import pytest

from util import random_mph_prims, relerr

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def test_slab_solver_single_gpu(gpu, oracle):
    import torch
    from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
    hs = gpu
    eos = (hs.Barton2009(), hs.Barton2009())
    nx, nsteps = 333, 12
    Ql, Qr = hs.initial_states(eos, 6)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    sol = SlabSolver(CudaKernels(eos, hs.MPH30, "cuda:0"), nx)
    sol.set_from_global(Q0)
    for _ in range(nsteps):
        sol.step(hs.HLL, 0.6, 1.0 / nx)
    Q = sol.gather()
    ref = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=8)
    assert relerr(Q, ref["Q"]) < 1e-9
    assert abs(sol.t[0] - ref["t"][0]) < 1e-11 * ref["t"][0] and sol.steps[0] == nsteps
    # same bits as the C-ABI context path
    with hs.Solver(eos, nx) as s2:
        s2.upload(Q0); s2.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps)
        assert np.array_equal(s2.download(), Q)
    # host-slab step (pinned buffers)
    hin = torch.as_tensor(Q0).pin_memory(); hout = torch.empty_like(hin).pin_memory()
    sol.step_host(hin, hout, hs.HLL, 0.6, 1.0 / nx)
    one = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, 1, nthreads=8)
    assert relerr(hout.numpy(), one["Q"]) < 1e-12


def test_step_host_c_abi(gpu, oracle):
    hs = gpu
    eos = hs.Barton2009()
    nx = 257
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 2)
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx, model=hs.SP13) as sol:
        Q1, dt = sol.step_host(Q0, None, "hll", 0.6, 1.0 / nx)
    one = oracle.run([oracle.barton2009()], oracle.SP13, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, 1)
    assert abs(dt[0] - one["dt"][0, 0]) < 1e-13 * dt[0]
    assert relerr(Q1, one["Q"]) < 1e-12


def test_ensemble_solver(gpu, oracle):
    from hyperelasticsolver_b200.slab import CudaKernels, EnsembleSolver
    hs = gpu
    rng = np.random.default_rng(9)
    eos = (hs.Barton2009(), hs.Barton2009())
    nprob, nx, nsteps = 5, 130, 10       # 130 cells: two tiles per problem with a nearly empty last tile
    Ql = hs.prim2cons_mph(eos, random_mph_prims(rng, nprob, spread=0.03)); Qr = hs.prim2cons_mph(eos, random_mph_prims(rng, nprob, spread=0.03))
    Q0 = np.stack([hs.initial_condition(Ql[i], Qr[i], nx) for i in range(nprob)])
    sol = EnsembleSolver(CudaKernels(eos, hs.MPH30, "cuda:0"), nx, nprob)
    sol.set_local(Q0)
    for _ in range(nsteps):
        sol.step(hs.HLL, 0.6, 1.0 / nx)
    ref = oracle.run(None, oracle.MPH30, oracle.HLL, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=8)
    assert relerr(sol.local().reshape(-1, 30), ref["Q"].reshape(-1, 30)) < 1e-9
    assert np.allclose(sol.t, ref["t"], rtol=1e-11)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _nccl_worker(rank, world, port, nx, nsteps, q, exchange="p2p"):
    os.environ["HS_EXCHANGE"] = exchange
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import hyperelasticsolver_b200 as hs
        from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
        eos = hs.Barton2009()
        Ql, Qr = hs.hyperelasticity.initial_states(eos, 1, device=rank)
        Q0 = hs.initial_condition(Ql, Qr, nx)
        sol = SlabSolver(CudaKernels(eos, hs.SP13, f"cuda:{rank}"), nx)
        sol.set_from_global(Q0)
        for _ in range(nsteps):
            sol.step(hs.HLL, 0.6, 1.0 / nx)
        Q = sol.gather()
        if rank == 0:
            with hs.Solver(eos, nx, model=hs.SP13, device=0) as s1:
                s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps)
                q.put((np.array_equal(s1.download(), Q), float(s1.t[0]), float(sol.t[0]), sol.exchange, getattr(sol, "_p2p_error", "")))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_slab_bit_identical(gpu, exchange):
    """2^n-independent check of BASELINE config 3's requirement: the slab-decomposed run is
    bit-identical to the single-GPU run (only an exact max crosses ranks) -- with the one-kernel
    peer-memory exchange and with the NCCL exchange."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, 5000, 25, q, exchange)) for r in range(2)]
    for p in procs: p.start()
    same, t1, t2, kind, err = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    assert same and t1 == t2
    assert kind == ("p2p-kernel" if exchange == "p2p" else "nccl"), f"exchange path {kind}: {err}"


def test_stateless_calls_keep_current_device(gpu):
    """The host-buffer entry points must not change the caller's current CUDA device (torch or a
    Julia allocator may be using another one)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    hs = gpu
    torch.cuda.set_device(1)
    x = torch.ones(4, device="cuda")
    eos = (hs.Barton2009(), hs.Barton2009())
    Ql, _ = hs.initial_states(eos, 6, device=0)
    with hs.Solver(eos, 64, device=0) as sol:
        sol.upload(hs.initial_condition(Ql, Ql, 64)); sol.step()
    assert torch.cuda.current_device() == 1
    assert float((x + 1).sum()) == 8.0
    torch.cuda.set_device(0)


@pytest.mark.parametrize("model", ["mph30", "sp13"])
def test_single_process_multi_device_context(gpu, oracle, model):
    """hs_create_multi: one host process (the Julia-driver situation), one grid slab-decomposed over the
    GPUs it can see, exchange by the peer-memory kernel.  Must be bit-identical to the one-GPU context."""
    import torch
    ndev = min(torch.cuda.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    hs = gpu
    if model == "mph30":
        eos = (hs.Barton2009(), hs.Barton2009()); hm = hs.MPH30
        Ql, Qr = hs.initial_states(eos, 6)
    else:
        eos = hs.Barton2009(); hm = hs.SP13
        Ql, Qr = hs.hyperelasticity.initial_states(eos, 2)
    nx, steps = 1003, 40          # odd size: unequal slabs
    Q0 = hs.initial_condition(Ql, Qr, nx)
    with hs.Solver(eos, nx, model=hm, device=0) as s1:
        s1.upload(Q0); h1 = s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=steps, record_dt=True)
        Q1 = s1.download(); lam1, eig1 = s1.wave_speeds(full=True)
    with hs.Solver(eos, nx, model=hm, devices=list(range(ndev))) as s2:
        s2.upload(Q0); h2 = s2.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=steps, record_dt=True)
        Q2 = s2.download(); lam2, eig2 = s2.wave_speeds(full=True)
        assert np.array_equal(h1, h2) and np.array_equal(Q1, Q2)
        assert np.array_equal(lam1, lam2) and np.array_equal(eig1, eig2)
        assert s2.steps[0] == steps and s2.t[0] == s1.t[0]
        # a second upload restarts cleanly (exchange sequence numbers keep increasing), t_end semantics hold
        s2.upload(Q0); s2.advance(float(h1[0, :7].sum()) * 0.999, "hll", 0.6, 1.0 / nx)
        assert s2.steps[0] == 7
        Qh, dt = s2.step_host(Q0, None, "hll", 0.6, 1.0 / nx)
    with hs.Solver(eos, nx, model=hm, device=0) as s1:
        Qh1, dt1 = s1.step_host(Q0, None, "hll", 0.6, 1.0 / nx)
    assert np.array_equal(Qh, Qh1) and dt[0] == dt1[0]


def test_single_process_multi_device_ensemble(gpu, oracle):
    """hs_create_multi with nprob > 1: problems shared out over the GPUs of the process, per-problem dt."""
    import torch
    ndev = min(torch.cuda.device_count(), 4)
    if ndev < 2:
        pytest.skip("needs >= 2 GPUs")
    hs = gpu
    rng = np.random.default_rng(17)
    eos = (hs.Barton2009(), hs.Barton2009())
    nprob, nx = 7, 90
    Ql = hs.prim2cons_mph(eos, random_mph_prims(rng, nprob, spread=0.03)); Qr = hs.prim2cons_mph(eos, random_mph_prims(rng, nprob, spread=0.03))
    Q0 = np.stack([hs.initial_condition(Ql[i], Qr[i], nx) for i in range(nprob)])
    t_end = 0.01
    out = []
    for devs in (None, list(range(ndev))):
        with hs.Solver(eos, nx, nprob=nprob, devices=devs) as sol:
            sol.upload(Q0)
            dt1 = sol.step("hll", 0.6, 1.0 / nx)
            hist = sol.advance(t_end, "hll", 0.6, 1.0 / nx, max_steps=300, record_dt=True)
            out.append((sol.download(), dt1.copy(), hist, sol.t.copy(), sol.steps.copy(), sol.wave_speeds()))
    for x, y in zip(out[0], out[1]):
        assert np.array_equal(x, y)
    assert len(set(out[0][4].tolist())) > 1          # problems really finish at different step counts


def test_exchange_gives_up_on_a_dead_peer(gpu):
    import subprocess
    env = dict(os.environ, HS_EXCHANGE_TIMEOUT_S="1")
    out = subprocess.run([sys.executable, os.path.join(HERE, "exchange_timeout_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "EXCHANGE-TIMEOUT-OK" in out.stdout, out.stdout[-1000:] + out.stderr[-3000:]


def _step_host_worker(rank, world, port, nx, nsteps, q, exchange="p2p"):
    """every rank keeps its slab (halo cells included) in pinned host memory and steps it with the chunk-pipelined
    SlabSolver.step_host; the gathered result must equal the single-GPU resident run bit for bit"""
    os.environ["HS_EXCHANGE"] = exchange
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import hyperelasticsolver_b200 as hs
        from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
        eos = hs.Barton2009()
        Ql, Qr = hs.hyperelasticity.initial_states(eos, 1, device=rank)
        x = (np.arange(nx) + 0.5) / nx
        w = (0.5 * (1 + np.tanh((x - 0.5) / 0.02)))[:, None]
        Q0 = np.ascontiguousarray((1 - w) * Ql[None, :] + w * Qr[None, :])
        sol = SlabSolver(CudaKernels(eos, hs.SP13, f"cuda:{rank}"), nx)
        a = torch.as_tensor(Q0[sol.lo_g:sol.hi_g].copy()).pin_memory()
        b = torch.empty_like(a).pin_memory()
        for _ in range(nsteps):
            sol.step_host(a, b, hs.HLL, 0.6, 1.0 / nx, chunk=2048)
            a, b = b, a
        mine = a.numpy()[sol.a - sol.lo_g: sol.b - sol.lo_g]
        parts = [None] * world
        dist.all_gather_object(parts, (sol.a, sol.b, mine, sol.pipelined_calls, sol.speculation_hits))
        if rank == 0:
            Q = np.empty_like(Q0)
            for (pa, pb, qq, _, _) in parts:
                Q[pa:pb] = qq
            with hs.Solver(eos, nx, model=hs.SP13, device=0) as s1:
                s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps)
                ref = s1.download()
            q.put((bool(np.array_equal(ref, Q)), float(np.abs(ref - Q).max()), [(p[3], p[4]) for p in parts], sol.exchange))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_two_gpu_pipelined_step_host(gpu, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    nsteps = 6
    procs = [ctx.Process(target=_step_host_worker, args=(r, 2, port, 40000, nsteps, q, exchange)) for r in range(2)]
    for p in procs: p.start()
    same, err, stats, kind = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120); assert p.exitcode == 0
    assert same, err
    assert stats == [(nsteps, nsteps - 1)] * 2, stats


def test_slab_solver_pipelined_step_host_single_rank(gpu):
    """world = 1: SlabSolver.step_host (pipelined, python host) == hs_step_host (pipelined, C) == the resident loop, bit for bit"""
    import torch
    hs = gpu
    from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
    nx, nsteps = 30000, 5
    for model in ("sp13", "mph30"):
        if model == "sp13":
            eos = hs.Barton2009(); Ql, Qr = hs.hyperelasticity.initial_states(eos, 1); hm = hs.SP13
        else:
            eos = (hs.Barton2009(), hs.Barton2009()); Ql, Qr = hs.initial_states(eos, 6); hm = hs.MPH30
        Q0 = hs.initial_condition(Ql, Qr, nx)
        sol = SlabSolver(CudaKernels(eos, hm, "cuda:0"), nx)
        a = torch.as_tensor(Q0.copy()).pin_memory(); b = torch.empty_like(a).pin_memory()
        for _ in range(nsteps):
            sol.step_host(a, b, hs.HLL, 0.6, 1.0 / nx, chunk=4096); a, b = b, a
        assert (sol.pipelined_calls, sol.speculation_hits) == (nsteps, nsteps - 1)
        with hs.Solver(eos, nx, model=hm) as s1:
            s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps)
            assert np.array_equal(s1.download(), a.numpy())
        # and the serial form
        sol2 = SlabSolver(CudaKernels(eos, hm, "cuda:0"), nx)
        a2 = torch.as_tensor(Q0.copy()).pin_memory(); b2 = torch.empty_like(a2).pin_memory()
        for _ in range(nsteps):
            sol2.step_host_serial(a2, b2, hs.HLL, 0.6, 1.0 / nx); a2, b2 = b2, a2
        assert np.array_equal(a2.numpy(), a.numpy())
