"""development aid: wall time of the shipped default run (BASELINE config 0) on the GPU"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hyperelasticsolver_b200 as H
eos = (H.Barton2009(), H.Barton2009())
Ql, Qr = H.initial_states(eos, 6)
Q0 = H.initial_condition(Ql, Qr, 1000)
with H.Solver(eos, 1000) as sol:
    sol.upload(Q0); sol.advance(0.06, "hll", 0.6, 1e-3)          # warm-up (module load, first launches)
    for rep in range(3):
        sol.upload(Q0)
        t0 = time.perf_counter(); sol.advance(0.06, "hll", 0.6, 1e-3); dt = time.perf_counter() - t0
        print(f"config 0 (nx=1000, T=0.06): {int(sol.steps[0])} steps in {dt*1e3:.1f} ms = {1000*sol.steps[0]/dt:.3e} cell-updates/s, {dt/sol.steps[0]*1e6:.1f} us/step")
