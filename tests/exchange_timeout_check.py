"""Run in a subprocess (HS_EXCHANGE_TIMEOUT_S=1) by tests/test_gpu_slab.py: a peer that never posts must
not hang the GPU -- the exchange kernel gives up after the timeout and flags the status word."""
import ctypes as C
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import torch
import hyperelasticsolver_b200 as hs
from hyperelasticsolver_b200 import _lib as L
from hyperelasticsolver_b200.slab import CudaKernels, scal_size

eos = hs.Barton2009()
k = CudaKernels(eos, hs.SP13, "cuda:0")
n = 64
prob = k.problem(n, 1)
Q = k.zeros(13, n); aux = k.zeros(k.naux, n); scal = k.zeros(scal_size(1))
mine = k.zeros(k.mailbox_doubles()); dead = k.zeros(k.mailbox_doubles())     # the "peer" never posts
t0 = time.perf_counter()
k.exchange_p2p(prob, Q, aux, scal[1:2], [mine.data_ptr(), dead.data_ptr()], 0, 2, 1, scal)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
status = int(scal[L.HS_SCAL_SLOTS:L.HS_SCAL_SLOTS + 1].view(torch.int32)[0].item())
assert status & 2, status
assert 0.8 < dt < 10.0, dt
print(f"EXCHANGE-TIMEOUT-OK {dt:.2f}s status={status}")
