"""development aid: small runs of every step-kernel flavour for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H
for model in (H.MPH30, H.SP13):
    for flux in ("hll", "lxf"):
        for gen in (False, True):
            if model == H.MPH30:
                eos = (H.Barton2009(), H.Barton2009(_beta=3.577, _gamma=2.088) if gen else H.Barton2009()); Ql, Qr = H.initial_states(eos, 6)
            else:
                eos = H.Barton2009(_beta=3.577, _gamma=2.088) if gen else H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 2)
            nx = 300
            Q0 = np.stack([H.initial_condition(Ql, Qr, nx)] * 2)
            with H.Solver(eos, nx, nprob=2, model=model) as s:
                s.upload(Q0); s.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=3); s.download(); s.wave_speeds(full=True)
eos = (H.Barton2009(), H.Barton2009()); Ql, Qr = H.initial_states(eos, 6)
Q = np.stack([Ql, Qr])
H.cons2prim_mph(eos, Q); H.flux_mph(eos, Q); H.noncons_flux(eos, Q); eg = H.get_eigvals(eos, Q); H.hll(eos, Ql, Qr, [eg[0], eg[1]]); H.lxf(eos, Ql, Qr, 10.0)
print("sanitize run done, launches:", H.lib().hs_kernel_launch_count())
