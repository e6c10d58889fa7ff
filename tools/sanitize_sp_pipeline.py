"""development aid: small runs of the single-phase TMA tile pipeline (k_step_sp) for compute-sanitizer: one grid (SINGLE) with
several tiles per block, odd cell counts (both column parities, thread-loaded last tile), ensembles (odd problem offsets,
problems that stop early), both fluxes, generic exponents."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H
os.environ["HS_SP_TILES"] = "3"
for gen in (False, True):
    eos = H.Barton2009(_beta=3.577, _gamma=2.088) if gen else H.Barton2009()
    Ql, Qr = H.hyperelasticity.initial_states(eos, 2)
    for flux in ("hll", "lxf"):
        for nx in (1000, 379, 130):
            with H.Solver(eos, nx, model=H.SP13) as s:
                s.upload(H.initial_condition(Ql, Qr, nx)); s.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=3); s.download()
        for nx, nprob in ((301, 5), (260, 4)):
            Q0 = np.stack([H.initial_condition(Ql, Qr, nx)] * nprob)
            with H.Solver(eos, nx, nprob=nprob, model=H.SP13) as s:
                s.upload(Q0); s.advance(3 * 0.6 / nx / 8.0, flux, 0.6, 1.0 / nx, max_steps=6); s.download()
print("sanitize run done, launches:", H.lib().hs_kernel_launch_count())
