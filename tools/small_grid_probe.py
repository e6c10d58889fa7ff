"""development aid: a few steps of the shipped nx = 1000 two-phase run (for an ncu launch list of the small-grid kernels)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyperelasticsolver_b200 as H
eos = (H.Barton2009(), H.Barton2009())
Ql, Qr = H.initial_states(eos, 6)
nx = int(os.environ.get("NX", "1000"))
Q0 = H.initial_condition(Ql, Qr, nx)
with H.Solver(eos, nx) as sol:
    sol.upload(Q0)
    sol.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=int(os.environ.get("STEPS", "30")))
    print(sol.steps)
