"""GPU, ONE device: the slab decomposition with ghost cells, as the multi-GPU drivers run it, on a single GPU.

`world` SlabSolver objects with faked ranks share the device; after every fused step their halo cells are exchanged
with hsd_halo pack / unpack + device copies and max(lambda) with an element-wise maximum (the NCCL exchange without
NCCL: tools/two_slabs_one_device.py).  With the odd interior cuts of slab_bounds every slab of an even-sized grid is an
even-sized array, so this drives k_step_sp<TM2D> WITH A GHOST MASK -- the kernel flavour every rank of `bench.py --gpus N`
runs -- on the 1-GPU test box, where the >= 2-GPU tests of test_gpu_slab.py are skipped.

Asserted: the gathered result is BIT-IDENTICAL to the single-domain run of the same library (only an exact max crosses
slabs), and within 1e-9 of the CPU oracle (main.jl:202-227)."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from util import relerr

pytestmark = pytest.mark.gpu


def _setup(hs, model, nx, tc=None):
    if model == hs.MPH30:
        eos = (hs.Barton2009(), hs.Barton2009())
        Ql, Qr = hs.initial_states(eos, 6 if tc is None else tc)
    else:
        eos = hs.Barton2009()
        Ql, Qr = hs.hyperelasticity.initial_states(eos, 1 if tc is None else tc)
    return eos, hs.initial_condition(Ql, Qr, nx)


@pytest.mark.parametrize("model,nx,world,nsteps,flux", [
    ("sp13", 20000, 2, 12, "hll"),     # even grid, odd cut: both slabs even-sized -> tensor-map tile copies + ghost mask
    ("sp13", 20001, 2, 12, "hll"),     # odd grid: one slab odd-sized -> row copies + ghost mask
    ("sp13", 40000, 4, 12, "hll"),     # interior slabs with two ghost cells
    ("sp13", 30000, 3, 10, "lxf"),
    ("sp13", 1 << 20, 8, 6, "hll"),    # many tiles per block in every slab (kper = 8 only above ~9500 tiles; here 1..2)
    ("sp13", 390, 3, 8, "hll"),        # slabs of about one tile; one slab spans a tile boundary
    ("mph30", 3000, 3, 6, "hll"),
    ("mph30", 4096, 2, 6, "lxf"),
    ("mph30", 8192, 4, 5, "hll"),
])
def test_slabs_on_one_device_bit_identical(gpu, oracle, model, nx, world, nsteps, flux):
    hs = gpu
    from hyperelasticsolver_b200 import _lib as L
    from hyperelasticsolver_b200.slab import CudaKernels
    from tools.two_slabs_one_device import run_slabs
    hmodel = hs.MPH30 if model == "mph30" else hs.SP13
    eos, Q0 = _setup(hs, hmodel, nx)
    with hs.Solver(eos, nx, model=hmodel) as s1:
        s1.upload(Q0)
        s1.advance(1e9, flux, 0.6, 1.0 / nx, max_steps=nsteps)
        ref = s1.download()
        t_ref = float(s1.t[0])
    kern = CudaKernels(eos, hmodel, "cuda:0")
    Q, t, nlocs = run_slabs(kern, Q0, world, nsteps, L.HLL if flux == "hll" else L.LXF)
    if nx % 2 == 0 and nx // world >= 64:
        assert all(n % 2 == 0 for n in nlocs), nlocs      # every slab takes the tensor-map flavour
    assert np.array_equal(Q, ref), f"max |d| = {np.abs(Q - ref).max():.3e}"
    assert t == t_ref
    if nx <= 40000:   # the oracle finishes these in seconds
        om = oracle.MPH30 if model == "mph30" else oracle.SP13
        oe = [oracle.barton2009()] * (2 if model == "mph30" else 1)
        r = oracle.run(oe, om, oracle.HLL if flux == "hll" else oracle.LXF, Q0, 0.6, 1.0 / nx, 1e9, nsteps, nthreads=oracle.hardware_threads())
        assert relerr(Q, r["Q"]) < 1e-9
        assert abs(t - r["t"][0]) <= 1e-12 * abs(r["t"][0])


def test_slabs_on_one_device_moving_data(gpu):
    """The same with data that differs in every cell (smooth volume-fraction profile + the Riemann jump), so that a halo cell
    that was not refreshed, or a ghost cell counted in max(lambda), cannot hide behind constant states."""
    hs = gpu
    from hyperelasticsolver_b200 import _lib as L
    from hyperelasticsolver_b200.slab import CudaKernels
    from tools.two_slabs_one_device import run_slabs
    eos = hs.Barton2009()
    nx = 6000
    Ql, Qr = hs.hyperelasticity.initial_states(eos, 1)
    x = (np.arange(nx) + 0.5) / nx
    w = (0.5 * (1 + np.tanh((x - 0.5) / 0.05)))[:, None]
    Q0 = (1 - w) * Ql[None, :] + w * Qr[None, :]
    Q0 *= (1.0 + 0.01 * np.sin(40 * np.pi * x))[:, None]   # scaling Q = rho (u, F, E) scales rho: still admissible
    with hs.Solver(eos, nx, model=hs.SP13) as s1:
        s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=25); ref = s1.download()
    kern = CudaKernels(eos, hs.SP13, "cuda:0")
    for world in (2, 3, 5):
        Q, t, nlocs = run_slabs(kern, Q0, world, 25, L.HLL)
        assert np.array_equal(Q, ref), (world, np.abs(Q - ref).max())
