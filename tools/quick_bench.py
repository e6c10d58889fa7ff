"""Quick device-resident timing of the fused step (development aid; bench.py is the contract)."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H

def run(model, logn, steps, flux="hll"):
    n = 1 << logn
    if model == H.MPH30:
        eos = (H.Barton2009(), H.Barton2009()); Ql, Qr = H.initial_states(eos, 6)
    else:
        eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
    Q0 = H.initial_condition(Ql, Qr, n)
    with H.Solver(eos, n, model=model) as sol:
        sol.upload(Q0)
        sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=3)
        t0 = time.perf_counter()
        sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=steps)
        dt = time.perf_counter() - t0
    cu = n * steps / dt
    nvar = 30 if model == H.MPH30 else 13
    print(json.dumps(dict(model="mph30" if model else "sp13", flux=flux, n=n, steps=steps, ms_per_step=1e3 * dt / steps,
                          cell_updates_per_s=cu, hbm_frac=cu * 2 * nvar * 8 / 6.45e12)))

if __name__ == "__main__":
    for logn in (20, 22):
        run(H.MPH30, logn, 10)
        run(H.SP13, logn, 10)
    run(H.MPH30, 22, 10, "lxf")
    run(H.SP13, 22, 10, "lxf")
