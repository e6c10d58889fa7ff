"""development aid: the fused step has no data-dependent shortcuts -- throughput on a Riemann grid
(piecewise constant) equals throughput on a grid where every cell differs (smooth random field)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hyperelasticsolver_b200 as H
from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver

def run(model, n, smooth, steps=10):
    if model == "mph30":
        eos = (H.Barton2009(), H.Barton2009()); hm = H.MPH30; nph = 2
    else:
        eos = H.Barton2009(); hm = H.SP13; nph = 1
    sol = SlabSolver(CudaKernels(eos, hm, "cuda:0"), n)
    dev = sol.k.device
    x = torch.arange(n, device=dev, dtype=torch.float64) / n
    # primitive fields: smooth, every cell different; or two constant states
    def field(base, amp, k):
        return base + (amp * torch.sin(2 * np.pi * k * x + base) if smooth else amp * torch.sign(x - 0.5))
    m = 4096
    # build a small table of conservative states through the library, then tile a smooth field by linear blending of states
    rng = np.random.default_rng(0)
    if model == "mph30":
        P = np.zeros((m, 30))
        for i in range(m):
            s = np.sin(2 * np.pi * i / m) if smooth else (1.0 if i >= m // 2 else -1.0)
            F = np.eye(3) + 0.02 * s * np.array([[1, 0.5, 0], [0.3, -1, 0.2], [0, 0.1, 0.7]])
            a1 = 0.5 + 0.3 * s
            for p, a in enumerate((a1, 1 - a1)):
                P[i, 15 * p:15 * p + 15] = [a, 8.9 / np.linalg.det(F), 0.3 * s, 0.1 * s, -0.2 * s, 5e-4 * (1 + s), *F.flatten(order="F")]
        Qt = H.prim2cons_mph(eos, P)
    else:
        P = np.zeros((m, 13))
        for i in range(m):
            s = np.sin(2 * np.pi * i / m) if smooth else (1.0 if i >= m // 2 else -1.0)
            F = np.eye(3) + 0.02 * s * np.array([[1, 0.5, 0], [0.3, -1, 0.2], [0, 0.1, 0.7]])
            P[i] = [0.3 * s, 0.1 * s, -0.2 * s, *F.flatten(order="C"), 5e-4 * (1 + s)]
        Qt = H.hyperelasticity.prim2cons(eos, P)
    Qt = torch.as_tensor(Qt, device=dev)
    pos = x * (m - 1)
    i0 = pos.floor().long().clamp(0, m - 2); w = (pos - i0) if smooth else torch.zeros_like(pos)
    if not smooth:
        i0 = torch.where(x < 0.5, torch.zeros_like(i0), torch.full_like(i0, m - 2)); w = torch.where(x < 0.5, torch.zeros_like(pos), torch.ones_like(pos))
    for v in range(sol.nvar):
        sol.Q[0][v] = Qt[i0, v] * (1 - w) + Qt[i0 + 1, v] * w
    sol.init_from_soa()
    for _ in range(3): sol.step(H.HLL, 0.6, 1.0 / n)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): sol.step(H.HLL, 0.6, 1.0 / n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    sol.check_status()
    distinct = int(torch.unique(sol.Q[sol.n & 1][sol.nvar - 1]).numel())
    print(json.dumps(dict(model=model, n=n, field="smooth (all cells differ)" if smooth else "two constant states", distinct_values=distinct,
                          gcups=round(n * steps / dt / 1e9, 3))))

for model, n in (("sp13", 1 << 23), ("mph30", 1 << 22)):
    for smooth in (False, True):
        run(model, n, smooth)
