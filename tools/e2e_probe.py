"""development aid (GPU box): what the host link gives -- H2D alone, D2H alone, both at once (pinned memory) -- and the
chunk-pipelined hs_step_host against that roof for a few chunk sizes."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import hyperelasticsolver_b200 as H

n = 1 << 24
nbytes = n * 13 * 8
hin = torch.empty(n, 13, dtype=torch.float64, pin_memory=True)
hout = torch.empty(n, 13, dtype=torch.float64, pin_memory=True)
d1 = torch.empty(n, 13, dtype=torch.float64, device="cuda")
d2 = torch.empty(n, 13, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timeit(f, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d1.copy_(hin, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hout.copy_(d2, non_blocking=True)
def both():
    h2d(); d2h()
def both_chunked(k=32):
    c = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d1[i * c:(i + 1) * c].copy_(hin[i * c:(i + 1) * c], non_blocking=True)
        with torch.cuda.stream(s2): hout[i * c:(i + 1) * c].copy_(d2[i * c:(i + 1) * c], non_blocking=True)
hin.zero_(); hout.zero_()
r = {"bytes": nbytes, "h2d_gbs": nbytes / timeit(h2d) / 1e9, "d2h_gbs": nbytes / timeit(d2h) / 1e9,
     "both_ms": 1e3 * timeit(both), "both_chunked32_ms": 1e3 * timeit(both_chunked)}
r["full_duplex_gbs_each"] = nbytes / (r["both_ms"] * 1e-3) / 1e9
print(json.dumps(r), flush=True)
del d1, d2

eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
gi = torch.arange(n)
hin.copy_(torch.where((gi < n / 2)[:, None], torch.as_tensor(Ql)[None, :], torch.as_tensor(Qr)[None, :]))
for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 21, 1 << 22):
    os.environ["HS_HOST_CHUNK"] = str(chunk)
    with H.Solver(eos, n, model=H.SP13) as sol:
        a, b = hin.numpy(), hout.numpy()
        for _ in range(3):
            sol.step_host(a, b, "hll", 0.6, 1.0 / n); a, b = b, a
        t0 = time.perf_counter()
        for _ in range(6):
            sol.step_host(a, b, "hll", 0.6, 1.0 / n); a, b = b, a
        ms = 1e3 * (time.perf_counter() - t0) / 6
        print(json.dumps({"chunk_cells": chunk, "step_host_ms": ms, "stats": sol.step_host_stats()}), flush=True)
os.environ["HS_HOST_PIPELINE"] = "0"
with H.Solver(eos, n, model=H.SP13) as sol:
    a, b = hin.numpy(), hout.numpy()
    sol.step_host(a, b, "hll", 0.6, 1.0 / n)
    t0 = time.perf_counter()
    for _ in range(4):
        sol.step_host(b, a, "hll", 0.6, 1.0 / n); a, b = b, a
    print(json.dumps({"plain_step_host_ms": 1e3 * (time.perf_counter() - t0) / 4}), flush=True)
