import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import hyperelasticsolver_b200 as hs
eos = hs.Barton2009(); Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
nx, ny = 300, 7
Q1 = hs.initial_condition(Ql, Qr, nx)
Q2 = np.ascontiguousarray(np.broadcast_to(Q1[None], (ny, nx, 13)))
with hs.Solver2D(eos, nx, ny, model=hs.SP13) as s2:
    s2.upload(Q2)
    print(s2.step("hll", 0.6, 1.0 / nx, 1e6))
