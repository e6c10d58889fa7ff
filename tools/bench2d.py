"""development aid: throughput of the dimension-split 2-D solver (not a tuned path: two extra transposes and CFL sweeps per step)"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H
for model, n in (("sp13", 4096), ("mph30", 2048)):
    if model == "sp13":
        eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1); hm = H.SP13
    else:
        eos = (H.Barton2009(), H.Barton2009()); Ql, Qr = H.initial_states(eos, 6); hm = H.MPH30
    row = H.initial_condition(Ql, Qr, n)
    Q = np.ascontiguousarray(np.broadcast_to(row[None], (n, n, row.shape[1])))
    with H.Solver2D(eos, n, n, model=hm) as s:
        s.upload(Q)
        s.advance(1e9, "hll", 0.6, 1.0 / n, 1.0 / n, max_steps=3)
        t0 = time.perf_counter()
        s.advance(1e9, "hll", 0.6, 1.0 / n, 1.0 / n, max_steps=10)
        dt = time.perf_counter() - t0
    print(json.dumps({"model": model, "grid": f"{n}x{n}", "ms_per_step": 1e3 * dt / 10, "gcups": n * n * 10 / dt / 1e9}))
