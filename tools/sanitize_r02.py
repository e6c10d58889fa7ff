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
# single-phase ensemble with several tiles per problem: the one-problem-per-block flavour of k_step_sp (even and odd problem lengths)
eos = H.Barton2009(); Ql, Qr = H.hyperelasticity.initial_states(eos, 2)
for nx, nprob in ((600, 3), (601, 4)):
    Q0 = np.stack([H.initial_condition(Ql, Qr, nx)] * nprob)
    with H.Solver(eos, nx, nprob=nprob, model=H.SP13) as s:
        s.upload(Q0); s.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=4); s.download()
# the persistent small-grid loop (cooperative launch) and the same steps launched one by one
eos2 = (H.Barton2009(), H.Barton2009()); Ql, Qr = H.initial_states(eos2, 6)
for loop in ("1", "0"):
    os.environ["HS_QP_LOOP"] = loop
    with H.Solver(eos2, 100) as s:
        s.upload(H.initial_condition(Ql, Qr, 100)); s.advance(1e9, "hll", 0.6, 0.01, max_steps=5); s.download()
os.environ.pop("HS_QP_LOOP")
# dimension-split 2-D solver: transposes with rotation, shared dt, both sweeps (two-phase: small-grid kernel; single-phase: row copies for odd ny)
for model, eos_, st in ((H.MPH30, eos2, H.initial_states(eos2, 6)), (H.SP13, eos, H.hyperelasticity.initial_states(eos, 1))):
    nx, ny = 70, 37
    row = H.initial_condition(st[0], st[1], nx)
    Q2 = np.ascontiguousarray(np.broadcast_to(row[None], (ny, nx, row.shape[1]))) * (1.0 + 0.01 * np.sin(np.arange(ny))[:, None, None])
    if model == H.MPH30:
        Q2[..., 0] = row[None, :, 0]; Q2[..., 15] = row[None, :, 15]
    with H.Solver2D(eos_, nx, ny, model=model) as s2:
        s2.upload(Q2); s2.step("hll", 0.6, 1.0 / nx, 1.0 / ny); s2.step("lxf", 0.6, 1.0 / nx, 1.0 / ny); s2.download()
print("sanitize run done, launches:", H.lib().hs_kernel_launch_count())
