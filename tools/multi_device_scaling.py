"""development aid: strong scaling of ONE grid through the single-process C-ABI context
(hs_create / hs_create_multi), i.e. what a Julia driver sees."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import hyperelasticsolver_b200 as hs

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 25
model = sys.argv[2] if len(sys.argv) > 2 else "sp13"
n = 1 << logn
if model == "sp13":
    eos = hs.Barton2009(); hm = hs.SP13; Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
else:
    eos = (hs.Barton2009(), hs.Barton2009()); hm = hs.MPH30; Ql, Qr = hs.initial_states(eos, 6)
Q0 = hs.initial_condition(Ql, Qr, n)
ndev = torch.cuda.device_count()
base = None
for k in [d for d in (1, 2, 4, 8) if d <= ndev]:
    with hs.Solver(eos, n, model=hm, devices=list(range(k))) as s:
        s.upload(Q0); s.advance(1e9, "hll", 0.6, 1.0 / n, max_steps=5)
        t0 = time.perf_counter(); s.advance(1e9, "hll", 0.6, 1.0 / n, max_steps=64); dt = time.perf_counter() - t0
        v = n * 64 / dt
        base = base or v
        print(json.dumps(dict(api="hs_create_multi (one process)", model=model, cells=n, devices=k, cell_updates_per_s=v, efficiency=v / (k * base))))
