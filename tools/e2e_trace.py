import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, numpy as np
import hyperelasticsolver_b200 as H
n = 1 << 24
hin = torch.empty(n, 13, dtype=torch.float64, pin_memory=True); hout = torch.empty(n, 13, dtype=torch.float64, pin_memory=True)
eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
gi = torch.arange(n)
hin.copy_(torch.where((gi < n / 2)[:, None], torch.as_tensor(Ql)[None, :], torch.as_tensor(Qr)[None, :]))
os.environ["HS_HOST_TRACE"] = "1"
with H.Solver(eos, n, model=H.SP13) as sol:
    a, b = hin.numpy(), hout.numpy()
    for _ in range(5):
        t0 = time.perf_counter(); sol.step_host(a, b, "hll", 0.6, 1.0 / n); a, b = b, a
        print("call ms", 1e3 * (time.perf_counter() - t0), flush=True)
