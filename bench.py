#!/usr/bin/env python
"""bench.py -- cell-updates/s (FP64) of the finite-volume time step, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of main.jl:204-227 (CFL dt + HLL update of every interior cell) over the
synthetic Riemann-problem grid.

Headline workload (when --workload is not given):
  N = 1 : BASELINE.json configs[1] -- single-phase (13-variable) 1-D Riemann problem, 2^24 cells, HLL, cfl 0.6
  N > 1 : BASELINE.json configs[3] -- the north-star multi-GPU case: single-phase 2^28 cells TOTAL, slab-decomposed
          over the N GPUs (strong scaling), halo cells + max(lambda) exchanged every step
and, in the same JSON line under "configs", the other BASELINE configurations measured in the same run (two-phase
2^24, 2^28 on one GPU / weak 2^24 per GPU, the 65,536 x 4,096 ensembles, the shipped nx = 1000 default run), each with
its own roofline block.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs in HBM), `e2e` = the same step through
the host-buffer API with the H2D/D2H copies of the whole state inside the timed region, `roofline` = the fused step
kernel against the measured HBM peak (and, for the two-phase kernel, against the measured FP64 issue peak, which is
its binding roof), `cpu_baseline` = the CPU oracle (a C++ restatement of the reference: Julia is not installed) on the
host cores, `parity_check` = a small slab-decomposed run of the same kernels compared bit for bit with the
single-domain run.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

WORKLOADS = {
    # name: (model, cells per GPU (weak) or total (strong), scaling, description)
    "sp13_2p24": dict(model="sp13", cells=1 << 24, scaling="weak", desc="single-phase 13-var 1-D Riemann problem (Hyperelasticity.jl test case 1), 2^24 cells per GPU, HLL"),
    "mph30_2p24": dict(model="mph30", cells=1 << 24, scaling="weak", desc="two-phase 30-var 1-D Riemann problem (HyperelasticityMPh.jl test case 6), 2^24 cells per GPU, HLL path-conservative"),
    "sp13_2p28": dict(model="sp13", cells=1 << 28, scaling="strong", desc="single-phase 1-D Riemann problem, 2^28 cells total, slab-decomposed, HLL"),
    "ensemble": dict(model="mph30", cells=4096, nprob=65536, scaling="strong", desc="65,536 independent two-phase Riemann problems x 4,096 cells, randomised states, per-problem dt"),
    "ensemble_sp": dict(model="sp13", cells=4096, nprob=65536, scaling="strong", desc="65,536 independent single-phase Riemann problems x 4,096 cells, randomised states, per-problem dt"),
}
METRIC = "cell-updates/sec (FP64)"
UNIT = "cell-updates/s"
TRAFFIC_JSON = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measure_fp64_peak():
    """FP64 FMA issue peak of this device, measured now with tools/fp64_peak (built by __graft_entry__.build()):
    the binding roof of the two-phase kernel.  Returns (DFMA/s, source)."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        out = subprocess.run([exe], capture_output=True, text=True, timeout=60).stdout
        best = max(json.loads(l)["dfma_per_s"] for l in out.splitlines() if l.startswith("{"))
        return float(best), "tools/fp64_peak, measured in this run"
    except Exception as e:   # noqa: BLE001
        return 1.708e13, f"profiles/r01_fp64_peak.jsonl (tools/fp64_peak did not run here: {e!r})"


def workload_config(name, world):
    """The `config` block of the JSON line: a pure function of (workload, N), so that the reference arm prints the same one."""
    wl = WORKLOADS[name]
    nvar = 30 if wl["model"] == "mph30" else 13
    if "nprob" in wl:
        total = wl["nprob"] * wl["cells"]
        per_gpu = (wl["nprob"] // world) * wl["cells"]
        par = f"ensemble partition x{world}, no collective"
    else:
        total = wl["cells"] * (world if wl["scaling"] == "weak" else 1)
        per_gpu = total // world
        par = f"slab x{world}, one halo cell per side + max(lambda) exchanged every step" if world > 1 else "one GPU"
    return {"workload": name, "description": wl["desc"], "flux": "hll", "cfl": 0.6, "cells_total": total, "cells_per_gpu": per_gpu,
            "parallelism": par,
            "l2": f"state {per_gpu * nvar * 8 / 1e9:.2f} GB per GPU per buffer >> 126 MB L2 (inputs larger than L2, no flush needed)"}


def numa_local_affinity(dev_index):
    """Bind this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers allocated next are
    NUMA-local to its PCIe root (first-touch placement).  Returns the node, or None if it cannot be determined."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(dev_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        bus = bus[-12:] if len(bus) > 12 else bus          # 00000000:1b:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:   # (virtualised hosts expose one node and no device affinity: nothing to bind)
            nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
            return f"not exposed by the host ({len(nodes)} NUMA node(s) visible)"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1 + 0.05)] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def riemann_states(H, model, device=0):
    if model == "mph30":
        eos = (H.Barton2009(), H.Barton2009())
        Ql, Qr = H.initial_states(eos, 6, device=device)
        return eos, H.MPH30, Ql, Qr
    eos = H.Barton2009()
    Ql, Qr = H.hyperelasticity.initial_states(eos, 1, device=device)
    return eos, H.SP13, Ql, Qr


def ensemble_states(H, model, p0, p1, seed=20261017, device=0):
    """BASELINE config 4 generator: PCG64(20261017), per problem and side
    alpha1~U[0.1,0.9], u~U[-1,1]^3, S~U[0,1e-3], F = I + 0.05 U[-1,1]^(3x3), nominal density 8.9."""
    rng = np.random.Generator(np.random.PCG64(seed))
    nprob = p1
    a = rng.uniform(0.1, 0.9, (nprob, 2)); u = rng.uniform(-1, 1, (nprob, 2, 3)); S = rng.uniform(0, 1e-3, (nprob, 2))
    F = np.eye(3)[None, None] + 0.05 * rng.uniform(-1, 1, (nprob, 2, 3, 3))
    a, u, S, F = a[p0:p1], u[p0:p1], S[p0:p1], F[p0:p1]
    det = np.linalg.det(F)
    if model == "mph30":
        P = np.zeros((p1 - p0, 2, 30))
        for ph, al in enumerate((a, 1 - a)):
            o = 15 * ph
            P[..., o] = al; P[..., o + 1] = 8.9 / det; P[..., o + 2:o + 5] = u; P[..., o + 5] = S
            P[..., o + 6:o + 15] = F.transpose(0, 1, 3, 2).reshape(p1 - p0, 2, 9)   # column-major
        eos = (H.Barton2009(), H.Barton2009())
        Q = H.prim2cons_mph(eos, P.reshape(-1, 30), device=device).reshape(p1 - p0, 2, 30)
        return eos, H.MPH30, Q
    P = np.concatenate([u, F.reshape(p1 - p0, 2, 9), S[..., None]], axis=-1)
    eos = H.Barton2009()
    Q = H.hyperelasticity.prim2cons(eos, P.reshape(-1, 13), device=device).reshape(p1 - p0, 2, 13)
    return eos, H.SP13, Q


def oracle_problem(model):
    import oracle as O
    from hyperelasticsolver_b200.testcases import mph_primitive_states, sp_primitive_states
    om = O.MPH30 if model == "mph30" else O.SP13
    eos = [O.barton2009()] * (2 if om == O.MPH30 else 1)
    Pl, Pr = mph_primitive_states(6) if om == O.MPH30 else sp_primitive_states(1)
    Qlr, _ = O.prim2cons(eos, om, np.stack([Pl, Pr]))
    return O, om, eos, Qlr


def cpu_oracle_rate(model, seconds, threads, literal=True):
    """The CPU restatement of main.jl on the host cores, on a bounded sample of the workload:
    a Riemann grid of `cells` cells around the interface for `steps` steps (~`seconds` of work)."""
    from hyperelasticsolver_b200.testcases import riemann_grid
    O, om, eos, Qlr = oracle_problem(model)
    cells = 4096
    t0 = time.perf_counter()
    O.run(eos, om, O.HLL, riemann_grid(Qlr[0], Qlr[1], cells), 0.6, 1.0 / cells, 1e9, 1, nthreads=threads, literal=literal)
    rate = cells / (time.perf_counter() - t0)
    steps = 4
    cells = int(min(1 << 20, max(4096, rate * seconds / steps)))
    Q0 = riemann_grid(Qlr[0], Qlr[1], cells)
    t0 = time.perf_counter()
    r = O.run(eos, om, O.HLL, Q0, 0.6, 1.0 / cells, 1e9, steps, nthreads=threads, literal=literal)
    dt = time.perf_counter() - t0
    assert r["status"] == 0
    return cells * steps / dt, f"{cells} cells x {steps} steps of the same Riemann problem, HLL, {'update_cell per cell (every face twice, as main.jl does)' if literal else 'each face once'}", dt


def default_workload(gpus):
    return "sp13_2p24" if gpus == 1 else "sp13_2p28"


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm on the host cores.  Julia is not
    installed in this image, so this is the C++ oracle restatement (kind 'port'), in literal mode
    (update_cell per cell, Threads.@threads -> std::thread over all host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from hyperelasticsolver_b200.testcases import riemann_grid
    name = args.workload or default_workload(args.gpus)
    wl = WORKLOADS[name]
    O, om, eos, Qlr = oracle_problem(wl["model"])
    threads = O.hardware_threads()
    # bounded sample: size the grid so that W+K steps take about two minutes at most
    probe = 2048
    t0 = time.perf_counter()
    O.run(eos, om, O.HLL, riemann_grid(Qlr[0], Qlr[1], probe), 0.6, 1.0 / probe, 1e9, 1, nthreads=threads, literal=True)
    rate = probe / (time.perf_counter() - t0)
    cells = int(min(1 << 20, max(2048, rate * args.ref_seconds / (args.steps + args.warmup))))
    Q = riemann_grid(Qlr[0], Qlr[1], cells)
    if args.warmup:
        Q = O.run(eos, om, O.HLL, Q, 0.6, 1.0 / cells, 1e9, args.warmup, nthreads=threads, literal=True)["Q"]
    t0 = time.perf_counter()
    r = O.run(eos, om, O.HLL, Q, 0.6, 1.0 / cells, 1e9, args.steps, nthreads=threads, literal=True)
    el = time.perf_counter() - t0
    v = cells * args.steps / el
    sample = (f"{cells} cells x {args.steps} steps: a bounded sample of the workload in `config` (same Riemann problem, flux and cfl; throughput per cell "
              f"does not depend on the grid size on the CPU), update_cell per cell as main.jl:221-226")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(name, args.gpus),
        "sample_cells": cells,
        "note": "CPU restatement of main.jl (C++ dual-number oracle), not Julia: julia is not installed in this image",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


# ------------------------------------------------------------------------------------------------
class Bench:
    """One process per GPU; all ranks walk through the same sequence of measurements."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import hyperelasticsolver_b200 as H
        from hyperelasticsolver_b200 import _lib as L
        self.torch, self.dist, self.H, self.L = torch, dist, H, L
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world != args.gpus and self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.peak, self.peak_src = measured_peaks()
        self.fp64_peak, self.fp64_src = (measure_fp64_peak() if self.rank == 0 else (1.708e13, ""))
        self.traffic = json.load(open(TRAFFIC_JSON)) if os.path.exists(TRAFFIC_JSON) else {}

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return int(x)
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.int64)
        self.dist.all_reduce(t)
        return int(t.item())

    # ---- build the synthetic state on the device, directly in SoA (not timed) ----------------------
    def setup(self, name):
        torch, H = self.torch, self.H
        from hyperelasticsolver_b200.slab import CudaKernels, EnsembleSolver, SlabSolver
        wl = WORKLOADS[name]
        model = wl["model"]
        nvar = 30 if model == "mph30" else 13
        world, rank, dev = self.world, self.rank, self.dev
        c = dict(name=name, wl=wl, model=model, nvar=nvar, ensemble="nprob" in wl)

        def fill(sol_, ql, qr, left_mask_cells, nprob_local=None):
            Q0 = sol_.Q[0]
            for v in range(nvar):
                if nprob_local is None:
                    Q0[v] = torch.where(left_mask_cells, ql[v], qr[v])
                else:
                    Q0[v].view(nprob_local, -1)[:] = torch.where(left_mask_cells[None, :], ql[:, v, None], qr[:, v, None])
            sol_.init_from_soa()

        if c["ensemble"]:
            nprob_g, ncells = wl["nprob"], wl["cells"]
            p0, p1 = nprob_g * rank // world, nprob_g * (rank + 1) // world
            eos, hmodel, Qlr = ensemble_states(H, model, p0, p1, device=self.local)
            kern = CudaKernels(eos, hmodel, dev)
            sol = EnsembleSolver(kern, ncells, nprob_g)
            left = torch.arange(ncells, device=dev) < ncells / 2
            Qlr_d = torch.as_tensor(Qlr, device=dev)
            fill(sol, Qlr_d[:, 0, :].contiguous(), Qlr_d[:, 1, :].contiguous(), left, sol.nprob)
            c.update(n_units=nprob_g * ncells, local_cells=sol.nprob * ncells, dx=1.0 / ncells, updated_local=sol.nprob * (ncells - 2), Qlr=Qlr,
                     exchange="none (independent problems)")
        else:
            eos, hmodel, Ql, Qr = riemann_states(H, model, device=self.local)
            n_global = wl["cells"] * (world if wl["scaling"] == "weak" else 1)
            kern = CudaKernels(eos, hmodel, dev)
            sol = SlabSolver(kern, n_global)
            gidx = torch.arange(sol.lo_g, sol.hi_g, device=dev)
            fill(sol, torch.as_tensor(Ql, device=dev), torch.as_tensor(Qr, device=dev), gidx < n_global / 2)
            del gidx
            ex = {"nccl": "NCCL send/recv + all-reduce", "p2p-kernel": "one peer-memory kernel over NVLink (hsd_exchange_p2p)"}[sol.exchange] if world > 1 else "nothing (1 GPU)"
            c.update(n_units=n_global, local_cells=sol.nloc, dx=1.0 / n_global, updated_local=sol.nloc - 2 - bin(sol.ghost_mask).count("1") + 0, Ql=Ql, Qr=Qr,
                     exchange=ex)
            # cells this rank UPDATES per launch: its owned cells minus the frozen physical boundary cells it holds
            c["updated_local"] = (sol.b - sol.a) - (1 if rank == 0 else 0) - (1 if rank == world - 1 else 0)
        c.update(eos=eos, hmodel=hmodel, kern=kern, sol=sol)
        torch.cuda.synchronize()
        return c

    # ---- device-resident throughput -----------------------------------------------------------------
    def resident(self, c, steps, warmup, sample_clocks=False):
        torch, L = self.torch, self.L
        sol, kern, flux, dx = c["sol"], c["kern"], L.HLL, c["dx"]
        for _ in range(warmup):
            sol.step(flux, 0.6, dx)
        torch.cuda.synchronize(); self.barrier()
        sampler = ClockSampler(self.local) if (sample_clocks and self.rank == 0) else None
        if sampler:
            sampler.start(); time.sleep(0.15)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = kern.launches()
        torch.cuda.synchronize(); self.barrier()
        tw0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            sol.step(flux, 0.6, dx, kernel_events=ev[i])
        e1.record()
        torch.cuda.synchronize()
        tw1 = time.perf_counter()
        self.barrier()
        launches = self.sum_over_ranks(kern.launches() - launches0)
        ms_total = self.max_over_ranks(e0.elapsed_time(e1))
        kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        clocks = sampler.stop(tw0, tw1) if sampler else None
        sol.check_status()
        return dict(value=c["n_units"] * steps / (ms_total * 1e-3), ms_per_step=ms_total / steps, kernel_ms=kernel_ms, launches=launches,
                    clocks=clocks, steps=steps, warmup=warmup)

    # ---- roofline of the fused step kernel ------------------------------------------------------------
    def roofline(self, c, m):
        nvar, model, name = c["nvar"], c["model"], c["name"]
        alg_bytes = 2 * nvar * 8 * c["updated_local"]
        achieved = alg_bytes / (m["kernel_ms"] * 1e-3) / 1e9
        tj = self.traffic.get(name) or {}
        tj_kernel = self.traffic.get("mph30_2p24" if model == "mph30" else "sp13_2p24") or {}
        traffic = tj.get("dram_bytes_per_launch") if self.world == 1 else None
        fp64_per_cu = tj_kernel.get("fp64_inst_per_cell_update")
        cu_per_s_kernel = c["updated_local"] / (m["kernel_ms"] * 1e-3)
        fp64 = None
        if fp64_per_cu:
            a = fp64_per_cu * cu_per_s_kernel
            fp64 = {"achieved": a / 1e12, "peak": self.fp64_peak / 1e12, "unit": "T FP64 instr/s (DFMA issue)", "frac": a / self.fp64_peak,
                    "fp64_inst_per_cell_update": fp64_per_cu, "peak_source": self.fp64_src, "count_source": tj_kernel.get("source")}
        hbm = {"achieved": achieved, "peak": self.peak, "unit": "GB/s", "frac": achieved / self.peak, "traffic": traffic}
        kernel = ("k_step_sp (TMA-fed tile pipeline: fused flux + update + next-step wave bounds)" if model == "sp13"
                  else "k_step (fused path-conservative flux + update + next-step wave bounds)")
        common = {"kernel": kernel, "kernel_ms": m["kernel_ms"], "algorithmic_bytes_per_cell_update": 2 * nvar * 8,
                  "cell_updates_per_launch": c["updated_local"], "peak_source": self.peak_src,
                  "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                  "fp64_pipe_pct_ncu": tj.get("fp64_pipe_pct_of_peak")}
        if model == "mph30":
            # binding roof of the two-phase kernel: FP64 issue; the HBM fraction is reported next to it
            r = {"bound": "fp64", **(fp64 or {"achieved": None, "peak": self.fp64_peak / 1e12, "unit": "T FP64 instr/s", "frac": None}),
                 "hbm": hbm, **common,
                 "note": "two-phase: ~18 FP64 instructions per algorithmic byte against a machine balance of ~2.6: the FP64 pipe is the roof; `hbm` is the north-star fraction by algorithmic bytes"}
            r["traffic"] = traffic
            return r
        return {"bound": "hbm", **hbm, **common, "fp64": fp64,
                "note": "single-phase: DRAM traffic (algorithmic 208 B + cached per-cell rows) and instruction issue are both near their limits (DESIGN.md)"}

    # ---- end to end through the host-buffer API ---------------------------------------------------------
    def e2e(self, c, e2e_steps):
        """every step: H2D of the whole (pinned) host state, CFL sweep, fused step, D2H of the new state.
        Grid: the workload itself when it is <= 2^24 cells per GPU, else a 2^24-cell-per-GPU sample of it
        (pinned host buffers of the 2^28 / ensemble workloads would not fit host memory twice)."""
        torch, H, L = self.torch, self.H, self.L
        from hyperelasticsolver_b200.slab import EnsembleSolver, SlabSolver
        wl, nvar, world, rank, kern = c["wl"], c["nvar"], self.world, self.rank, c["kern"]
        flux = L.HLL
        numa = numa_local_affinity(self.local)      # pinned buffers below land on the GPU's own NUMA node
        if c["ensemble"]:
            nprob_e = min(wl["nprob"], (1 << 24) // wl["cells"] * world)
            pe0, pe1 = nprob_e * rank // world, nprob_e * (rank + 1) // world
            np_loc = pe1 - pe0
            s2 = H.Solver(c["eos"], wl["cells"], nprob=np_loc, model=c["hmodel"], device=self.local)   # every rank steps its own share: no exchange
            e_cells_local, e_units = np_loc * wl["cells"], nprob_e * wl["cells"]
            host_in = torch.empty(e_cells_local, nvar, dtype=torch.float64, pin_memory=True)
            host_out = torch.empty_like(host_in).pin_memory()
            left_h = (torch.arange(wl["cells"]) < wl["cells"] / 2)[None, :, None]
            q = torch.as_tensor(c["Qlr"][:np_loc])
            host_in.view(np_loc, wl["cells"], nvar)[:] = torch.where(left_h, q[:, None, 0, :], q[:, None, 1, :])
            step_host = lambda a, b: s2.step_host(a.numpy().reshape(np_loc, wl["cells"], nvar), b.numpy().reshape(np_loc, wl["cells"], nvar), "hll", 0.6, c["dx"])
            api = ("hs_step_host (C ABI) on this rank's share of the problems: chunk-pipelined by groups of whole problems "
                   "(H2D || CFL sweep + fused step || D2H, per-problem speculative dt)")
        else:
            e_cells = min(wl["cells"] if wl["scaling"] == "weak" else wl["cells"] // world, 1 << 24)
            e_units = e_cells * world
            if world == 1:
                s2 = H.Solver(c["eos"], e_cells, model=c["hmodel"], device=self.local)
                e_cells_local, lo_g = e_cells, 0
                step_host = lambda a, b: s2.step_host(a.numpy(), b.numpy(), "hll", 0.6, 1.0 / e_units)
                api = "hs_step_host (C ABI): upload Q0 + CFL sweep + fused step + download Q1, every step, chunk-pipelined (H2D || kernels || D2H)"
            else:
                sol = SlabSolver(kern, e_units)
                e_cells_local, lo_g = sol.nloc, sol.lo_g
                step_host = lambda a, b: sol.step_host(a, b, flux, 0.6, 1.0 / e_units)
                api = ("SlabSolver.step_host: pinned host slab -> device, CFL sweep, fused step, exchange, device -> host, every step, "
                       "chunk-pipelined over three streams per rank (speculative dt, confirmed by an all-reduce of the sweep's max(lambda))")
            host_in = torch.empty(e_cells_local, nvar, dtype=torch.float64, pin_memory=True)
            host_out = torch.empty_like(host_in).pin_memory()
            gi = torch.arange(lo_g, lo_g + e_cells_local)
            host_in.copy_(torch.where((gi < e_units / 2)[:, None], torch.as_tensor(c["Ql"])[None, :], torch.as_tensor(c["Qr"])[None, :]))
        step_host(host_in, host_out)                     # warm-up (allocations, first touch)
        step_host(host_out, host_in)
        step_host(host_in, host_out)
        torch.cuda.synchronize(); self.barrier()
        l0 = kern.launches()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step_host(host_out if i % 2 == 0 else host_in, host_in if i % 2 == 0 else host_out)
        torch.cuda.synchronize()
        e2e_t = self.max_over_ranks(time.perf_counter() - t0)
        self.barrier()
        nbytes = e_units * nvar * 8
        return {"value": e_units * e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                "steps": e2e_steps, "ms_per_step": 1e3 * e2e_t / e2e_steps, "cells_total": e_units, "api": api,
                "gpu_launches_per_rank": kern.launches() - l0, "host_buffers": "pinned (torch), NUMA node of the GPU: " + str(numa),
                "link_roof_note": "PCIe Gen5 x16 measured on this pool: 55.5 GB/s one direction alone, 48 GB/s each with both directions busy (profiles/r02_e2e_link_probe.log)"}

    # ---- correctness carried by the bench line -----------------------------------------------------------
    def parity_check(self):
        """A small slab-decomposed run (2^16 cells, 8 steps, both models) through the same kernels and the same exchange as the
        timed loop, gathered and compared BIT FOR BIT with rank 0's single-domain run of the same library (only an exact max
        crosses slabs, so any difference is a bug).  N = 1: two / three slabs on the one device (tools/two_slabs_one_device.py)."""
        torch, H, L = self.torch, self.H, self.L
        from hyperelasticsolver_b200.slab import CudaKernels, SlabSolver
        out = {"cells": 1 << 16, "steps": 8, "bit_identical": True, "max_rel_vs_single": 0.0, "cases": []}
        for model in ("sp13", "mph30"):
            eos, hmodel, Ql, Qr = riemann_states(H, model, device=self.local)
            nx, nsteps = 1 << 16, 8
            Q0 = H.initial_condition(Ql, Qr, nx)
            kern = CudaKernels(eos, hmodel, self.dev)
            if self.world > 1:
                sol = SlabSolver(kern, nx)
                sol.set_from_global(Q0)
                for _ in range(nsteps):
                    sol.step(L.HLL, 0.6, 1.0 / nx)
                sol.check_status()
                Q = sol.gather()
                how = f"{self.world} ranks, exchange: {sol.exchange}"
                del sol
            else:
                from tools.two_slabs_one_device import run_slabs
                Q, _, nlocs = run_slabs(kern, Q0, 3, nsteps, L.HLL)
                how = f"3 slabs on one device (local sizes {nlocs}), hsd_halo + device copies"
            if self.rank == 0:
                with H.Solver(eos, nx, model=hmodel, device=self.local) as s1:
                    s1.upload(Q0); s1.advance(1e9, "hll", 0.6, 1.0 / nx, max_steps=nsteps); ref = s1.download()
                same = bool(np.array_equal(Q, ref))
                rel = float((np.abs(Q - ref).max(axis=0) / np.maximum(np.abs(ref).max(axis=0), 1e-300)).max())
                out["bit_identical"] &= same
                out["max_rel_vs_single"] = max(out["max_rel_vs_single"], rel)
                out["cases"].append({"model": model, "how": how, "bit_identical": same})
            self.barrier()
        return out

    def config0(self):
        """BASELINE configs[0]: the run main.jl ships (two-phase test case 6, nx = 1000, cfl 0.6, T = 0.06, HLL: 641 steps) through
        hs_advance, wall time on this GPU (grid too small to fill it: latency-bound)."""
        H = self.H
        eos = (H.Barton2009(), H.Barton2009())
        Ql, Qr = H.initial_states(eos, 6, device=self.local)
        nx = 1000
        Q0 = H.initial_condition(Ql, Qr, nx)
        best, steps = None, 0
        with H.Solver(eos, nx, device=self.local) as sol:
            for _ in range(3):
                sol.upload(Q0)
                t0 = time.perf_counter()
                sol.advance(0.06, "hll", 0.6, 1.0 / nx)
                el = time.perf_counter() - t0
                best = el if best is None else min(best, el)
                steps = int(sol.steps[0]); t_end = float(sol.t[0])
        return {"workload": "main.jl default: two-phase test case 6, nx = 1000, cfl 0.6, T = 0.06, HLL", "steps": steps, "t_end": t_end,
                "gpu_wall_ms": 1e3 * best, "us_per_step": 1e6 * best / max(steps, 1), "value": nx * steps / best, "unit": UNIT,
                "api": "hs_advance (device-resident loop)"}

    def release(self, c):
        c.pop("sol", None); c.pop("kern", None)
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()

    def sub_result(self, name, steps, warmup):
        c = self.setup(name)
        m = self.resident(c, steps, warmup)
        r = {"value": m["value"], "unit": UNIT, "ms_per_step": m["ms_per_step"], "steps": steps, "warmup": warmup, "scaling": c["wl"]["scaling"],
             "config": workload_config(name, self.world), "exchange": c["exchange"], "roofline": self.roofline(c, m), "gpu_launches": m["launches"]}
        if name in ("mph30_2p24", "ensemble_sp") and self.world == 1:
            # end to end with the state in host memory: the model main.jl ships (2 x 4.03 GB over the link per step) / a 2^24-cell sample of the ensemble
            c.pop("sol")
            self.torch.cuda.empty_cache()
            r["e2e"] = self.e2e(c, 3)
        self.release(c)
        return r

    def run(self):
        args = self.args
        name = args.workload or default_workload(self.world)
        c = self.setup(name)
        m = self.resident(c, args.steps, args.warmup, sample_clocks=True)
        roofline = self.roofline(c, m)
        exchange = c["exchange"]
        c.pop("sol")
        self.torch.cuda.empty_cache()
        e2e = self.e2e(c, max(1, min(args.steps, args.e2e_steps)))
        self.release(c)
        parity = self.parity_check() if not args.no_parity_check else None

        configs = {}
        if args.workload is None and not args.no_subconfigs:
            sub_steps, sub_warm = min(args.steps, 10), 3
            others = ["mph30_2p24", "sp13_2p28", "ensemble_sp", "ensemble"] if self.world == 1 else ["sp13_2p24", "mph30_2p24", "ensemble_sp", "ensemble"]
            for o in others:
                configs[o] = self.sub_result(o, sub_steps if o != "ensemble" else min(sub_steps, 5), sub_warm)
            if self.rank == 0:
                configs["config0_default_run"] = self.config0()
            self.barrier()

        cpu = None
        if self.rank == 0 and self.world == 1 and not args.no_cpu_baseline:
            import oracle as O
            threads = O.hardware_threads()
            v, sample, el = cpu_oracle_rate(c["model"], args.cpu_seconds, threads, literal=True)
            v1, sample1, el1 = cpu_oracle_rate(c["model"], min(4.0, args.cpu_seconds), 1, literal=True)   # also single-thread
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "seconds": el,
                   "single_thread": {"value": v1, "sample": sample1, "seconds": el1},
                   "note": "C++ restatement of main.jl (dual-number AD like ForwardDiff); Julia itself is not installed in this image"}
            if "config0_default_run" in configs:   # the shipped default run on the host cores, same run
                from hyperelasticsolver_b200.testcases import riemann_grid
                Oo, om, oe, Qlr = oracle_problem("mph30")
                t0 = time.perf_counter()
                r = Oo.run(oe, om, Oo.HLL, riemann_grid(Qlr[0], Qlr[1], 1000), 0.6, 1e-3, 0.06, 40, nthreads=threads, literal=True)
                el0 = time.perf_counter() - t0
                configs["config0_default_run"]["cpu_port_ms_per_step"] = 1e3 * el0 / 40
                configs["config0_default_run"]["cpu_port_sample"] = f"first 40 of the 641 steps on {threads} host threads"

        if self.rank == 0:
            out = {
                "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": self.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": c["wl"]["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(name, self.world), "exchange": exchange,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": m["launches"], "clocks": m["clocks"],
                "parity_check": parity, "configs": configs,
            }
            print(json.dumps(out))
        if self.world > 1:
            self.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: sp13_2p24 on one GPU, sp13_2p28 (strong) on several, plus the other BASELINE configs as sub-results")
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="--impl reference: CPU work the bounded sample is sized for")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-subconfigs", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        Bench(args).run()


if __name__ == "__main__":
    main()
