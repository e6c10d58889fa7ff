"""Development aid: device-resident timing of the single-phase step for the plain kernel (HS_SP_TMA=0) and the TMA
tile pipeline at several tiles-per-block settings (bench.py is the contract; this is the tuning loop)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H

def run(logn, steps, flux, env):
    for k in ("HS_SP_TMA", "HS_SP_TILES", "HS_SP_SINGLE", "HS_SP_TMA2D"):
        os.environ.pop(k, None)
    os.environ.update(env)
    n = 1 << logn
    eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 1)
    Q0 = H.initial_condition(Ql, Qr, n)
    with H.Solver(eos, n, model=H.SP13) as sol:
        sol.upload(Q0)
        sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=3)
        best = 0.0
        for _ in range(3):
            t0 = time.perf_counter()
            sol.advance(1e9, flux, 0.6, 1.0 / n, max_steps=steps)
            best = max(best, n * steps / (time.perf_counter() - t0))
    print(json.dumps(dict(n=n, flux=flux, env=env, gcups=round(best / 1e9, 3), hbm_frac=round(best * 208 / 6.45e12, 4))), flush=True)

if __name__ == "__main__":
    logn = int(sys.argv[1]) if len(sys.argv) > 1 else 23
    settings = [{"HS_SP_TMA": "0"}] + [{"HS_SP_TILES": str(k)} for k in (1, 2, 4, 8, 16, 32)]
    for env in settings:
        run(logn, 20, "hll", env)
    run(logn, 20, "lxf", {"HS_SP_TMA": "0"})
    run(logn, 20, "lxf", {})
