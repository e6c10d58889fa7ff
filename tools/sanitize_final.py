"""development aid: compute-sanitizer run of the kernels added last in the round -- the tensor-map flavour of the single-phase
tile pipeline (even strides; odd ones take the row copies) and the Hank2016 batches (full blocks and a ragged one)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hyperelasticsolver_b200 as H
from hyperelasticsolver_b200 import equations_of_state as E
os.environ["HS_SP_TILES"] = "3"
eos = H.Barton2009()
Ql, Qr = H.hyperelasticity.initial_states(eos, 2)
for flux in ("hll", "lxf"):
    for nx in (1000, 130, 379):
        with H.Solver(eos, nx, model=H.SP13) as s:
            s.upload(H.initial_condition(Ql, Qr, nx)); s.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=3); s.download()
    for nx, nprob in ((260, 4), (301, 5)):
        Q0 = np.stack([H.initial_condition(Ql, Qr, nx)] * nprob)
        with H.Solver(eos, nx, nprob=nprob, model=H.SP13) as s:
            s.upload(Q0); s.advance(3 * 0.6 / nx / 8.0, flux, 0.6, 1.0 / nx, max_steps=6); s.download()
rng = np.random.default_rng(0)
for n in (1, 128, 300):
    A = np.eye(3)[None] + 0.1 * rng.uniform(-1, 1, (n, 3, 3))
    a9 = A.transpose(0, 2, 1).reshape(n, 9).copy()
    g9 = np.einsum("nki,nkj->nij", A, A).transpose(0, 2, 1).reshape(n, 9).copy()
    den = np.full(n, 2.7); pres = np.full(n, 1e9)
    e = E.energy(E.eos_hank2016, den, pres, g9)
    inv3 = np.stack([g9[:, 0] + g9[:, 4] + g9[:, 8], np.ones(n), np.ones(n)], axis=1)
    E.pressure(E.eos_hank2016, den, e, inv3)
    E.stress(E.eos_hank2016, den, pres, a9)
print("sanitize run done, launches:", H.lib().hs_kernel_launch_count())
