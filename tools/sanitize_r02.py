"""development aid: the code paths added in round 2, small, for compute-sanitizer (memcheck / racecheck / synccheck):
k_step_qp (quadrature-parallel small-grid step: its shared-memory phases), the window launches of the chunk-pipelined
hs_step_host (ghost masks, tensor-map copies on windows), graph replay in hs_advance, the persistent small-grid loop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H

os.environ["HS_HOST_CHUNK"] = "1024"
for model in (H.MPH30, H.SP13):
    for flux in ("hll", "lxf"):
        for gen in (False, True):
            if model == H.MPH30:
                eos = (H.Barton2009(), H.Barton2009(_beta=3.577, _gamma=2.088) if gen else H.Barton2009()); Ql, Qr = H.initial_states(eos, 6)
            else:
                eos = H.Barton2009(_beta=3.577, _gamma=2.088) if gen else H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 2)
            for nx, nprob in ((45, 1), (300, 2)):                       # k_step_qp for the two-phase model (tiles of 14 cells, partial last tile)
                Q0 = np.stack([H.initial_condition(Ql, Qr, nx)] * nprob)
                with H.Solver(eos, nx, nprob=nprob, model=model) as s:
                    s.upload(Q0); s.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=14, record_dt=True); s.download()   # 12 steps from the graph + 2
            nx = 4200                                                    # pipelined host step: 4-5 chunks, ghost windows
            Q0 = H.initial_condition(Ql, Qr, nx)
            with H.Solver(eos, nx, model=model) as s:
                Q1, _ = s.step_host(Q0, None, flux); Q2, _ = s.step_host(Q1, None, flux)
                assert s.step_host_stats() == (2, 1)
print("sanitize run done, launches:", H.lib().hs_kernel_launch_count())
