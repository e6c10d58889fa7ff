"""development aid: throughput of the generic-exponent (exp/log) instantiation = heterogeneous EoS of main.jl:133"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyperelasticsolver_b200 as H
n = 1 << 22
for name, eos in (("default (1,3,2)", (H.Barton2009(), H.Barton2009())),
                  ("main.jl:133 copper pair", (H.Barton2009(), H.Barton2009(_rho0=8.93, _c0=6.22, _cv=9.0e-4, _t0=300, _b0=3.16, _alpha=1, _beta=3.577, _gamma=2.088)))):
    Ql, Qr = H.initial_states(eos, 6)
    Q0 = H.initial_condition(Ql, Qr, n)
    with H.Solver(eos, n) as sol:
        sol.upload(Q0); sol.advance(1e9, "hll", 0.6, 1.0 / n, max_steps=3)
        t0 = time.perf_counter(); sol.advance(1e9, "hll", 0.6, 1.0 / n, max_steps=10); dt = time.perf_counter() - t0
    print(json.dumps(dict(eos=name, model="mph30", n=n, gcups=round(n * 10 / dt / 1e9, 3))))
