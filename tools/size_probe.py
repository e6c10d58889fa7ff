"""development aid: single-phase step throughput against the grid size (power-of-two row pitch vs a slightly padded one)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hyperelasticsolver_b200 as H
from hyperelasticsolver_b200 import _lib as L
from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
dev = torch.device("cuda", 0)
kern = CudaKernels(eos, H.SP13, dev)
ql, qr = torch.as_tensor(Ql, device=dev), torch.as_tensor(Qr, device=dev)
for n in [int(x) for x in sys.argv[1:]]:
    sol = SlabSolver(kern, n)
    left = torch.arange(n, device=dev) < n / 2
    for v in range(13):
        sol.Q[0][v] = torch.where(left, ql[v], qr[v])
    del left
    sol.init_from_soa()
    for _ in range(3):
        sol.step(L.HLL, 0.6, 1.0 / n)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            sol.step(L.HLL, 0.6, 1.0 / n)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    print(json.dumps({"n": n, "ms": best, "gcups": n / best / 1e6, "frac": n * 208 / (best * 1e-3) / 6.551e12}), flush=True)
    del sol
    torch.cuda.empty_cache()
