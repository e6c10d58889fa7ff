#!/usr/bin/env python
"""bench.py -- cell-updates/s (FP64) of the finite-volume time step, BASELINE.json's metric.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of main.jl:204-227 (CFL dt + HLL update of every interior cell) over the
synthetic Riemann-problem grid.  Default workload = BASELINE.json configs[1]: single-phase
(13-variable) 1-D Riemann problem, 2^24 cells per GPU, HLL flux, cfl 0.6 (weak scaling: N GPUs
carry one global grid of N*2^24 cells, slab-decomposed with halo exchange + allreduce(max)).

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs in HBM), `e2e` =
the same step through the host-buffer API with the H2D/D2H copies of the whole state inside the
timed region, `roofline` = the fused step kernel against the measured HBM peak, `cpu_baseline`
= the CPU oracle (a C++ restatement of the reference: Julia is not installed) on the host cores.
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
    # name: (model, log2 cells per GPU or total, scaling, description)
    "sp13_2p24": dict(model="sp13", cells=1 << 24, scaling="weak", desc="single-phase 13-var 1-D Riemann problem (Hyperelasticity.jl test case 1), 2^24 cells per GPU, HLL"),
    "mph30_2p24": dict(model="mph30", cells=1 << 24, scaling="weak", desc="two-phase 30-var 1-D Riemann problem (HyperelasticityMPh.jl test case 6), 2^24 cells per GPU, HLL path-conservative"),
    "sp13_2p28": dict(model="sp13", cells=1 << 28, scaling="strong", desc="single-phase 1-D Riemann problem, 2^28 cells total, slab-decomposed, HLL"),
    "ensemble": dict(model="mph30", cells=4096, nprob=65536, scaling="strong", desc="65,536 independent two-phase Riemann problems x 4,096 cells, randomised states, per-problem dt"),
    "ensemble_sp": dict(model="sp13", cells=4096, nprob=65536, scaling="strong", desc="65,536 independent single-phase Riemann problems x 4,096 cells, randomised states, per-problem dt"),
}
METRIC = "cell-updates/sec (FP64)"
UNIT = "cell-updates/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def cpu_oracle_rate(model, seconds, threads, literal=True):
    """The CPU restatement of main.jl on the host cores, on a bounded sample of the workload:
    a Riemann grid of `cells` cells around the interface for `steps` steps (~`seconds` of work)."""
    import oracle as O
    om = O.MPH30 if model == "mph30" else O.SP13
    from hyperelasticsolver_b200.testcases import mph_primitive_states, riemann_grid, sp_primitive_states
    eos = [O.barton2009()] * (2 if om == O.MPH30 else 1)
    Pl, Pr = mph_primitive_states(6) if om == O.MPH30 else sp_primitive_states(1)
    Qlr, _ = O.prim2cons(eos, om, np.stack([Pl, Pr]))
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


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm on the host cores.  Julia is not
    installed in this image, so this is the C++ oracle restatement (kind 'port'), in literal mode
    (update_cell per cell, Threads.@threads -> std::thread over all host cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    wl = WORKLOADS[args.workload]
    threads = O.hardware_threads()
    om = O.MPH30 if wl["model"] == "mph30" else O.SP13
    from hyperelasticsolver_b200.testcases import mph_primitive_states, riemann_grid, sp_primitive_states
    eos = [O.barton2009()] * (2 if om == O.MPH30 else 1)
    Pl, Pr = mph_primitive_states(6) if om == O.MPH30 else sp_primitive_states(1)
    Qlr, _ = O.prim2cons(eos, om, np.stack([Pl, Pr]))
    # bounded sample: size the grid so that W+K steps take about two minutes at most
    probe = 2048
    t0 = time.perf_counter()
    O.run(eos, om, O.HLL, riemann_grid(Qlr[0], Qlr[1], probe), 0.6, 1.0 / probe, 1e9, 1, nthreads=threads, literal=True)
    rate = probe / (time.perf_counter() - t0)
    cells = int(min(1 << 20, max(2048, rate * 90.0 / (args.steps + args.warmup))))
    Q = riemann_grid(Qlr[0], Qlr[1], cells)
    if args.warmup:
        Q = O.run(eos, om, O.HLL, Q, 0.6, 1.0 / cells, 1e9, args.warmup, nthreads=threads, literal=True)["Q"]
    t0 = time.perf_counter()
    r = O.run(eos, om, O.HLL, Q, 0.6, 1.0 / cells, 1e9, args.steps, nthreads=threads, literal=True)
    el = time.perf_counter() - t0
    v = cells * args.steps / el
    sample = f"{cells} cells x {args.steps} steps (bounded sample of the {wl['desc']}), update_cell per cell as main.jl:221-226"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": args.workload, "description": wl["desc"], "flux": "hll", "cfl": 0.6,
                                        "note": "CPU restatement of main.jl (C++ dual-number oracle), not Julia: julia is not installed in this image"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import hyperelasticsolver_b200 as H
    from hyperelasticsolver_b200 import _lib as L
    from hyperelasticsolver_b200.slab import CudaKernels, EnsembleSolver, SlabSolver, slab_bounds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = WORKLOADS[args.workload]
    model = wl["model"]
    nvar = 30 if model == "mph30" else 13
    flux = L.HLL
    ensemble = "nprob" in wl

    # ---- build the synthetic state on the device, directly in SoA (not timed) --------------------
    def fill_riemann_soa(sol_, ql, qr, left_mask_cells, nprob_local=None):
        """Q[0][v] = left ? ql[v] : qr[v]; ql/qr: (nvar,) or (nprob_local, nvar) device tensors."""
        Q0 = sol_.Q[0]
        for v in range(nvar):
            if nprob_local is None:
                Q0[v] = torch.where(left_mask_cells, ql[v], qr[v])
            else:
                Q0[v].view(nprob_local, -1)[:] = torch.where(left_mask_cells[None, :], ql[:, v, None], qr[:, v, None])
        sol_.init_from_soa()

    exchange_kind = "none"
    if ensemble:
        nprob_g, ncells = wl["nprob"], wl["cells"]
        p0, p1 = nprob_g * rank // world, nprob_g * (rank + 1) // world
        eos, hmodel, Qlr = ensemble_states(H, model, p0, p1, device=local)
        kern = CudaKernels(eos, hmodel, dev)
        sol = EnsembleSolver(kern, ncells, nprob_g)
        left = torch.arange(ncells, device=dev) < ncells / 2
        Qlr_d = torch.as_tensor(Qlr, device=dev)
        fill_riemann_soa(sol, Qlr_d[:, 0, :].contiguous(), Qlr_d[:, 1, :].contiguous(), left, sol.nprob)
        n_units = nprob_g * ncells
        local_cells = sol.nprob * ncells
        dx = 1.0 / ncells
        updated_local = sol.nprob * (ncells - 2)
    else:
        eos, hmodel, Ql, Qr = riemann_states(H, model, device=local)
        n_global = wl["cells"] * (world if wl["scaling"] == "weak" else 1)
        kern = CudaKernels(eos, hmodel, dev)
        sol = SlabSolver(kern, n_global)
        gidx = torch.arange(sol.lo_g, sol.hi_g, device=dev)
        fill_riemann_soa(sol, torch.as_tensor(Ql, device=dev), torch.as_tensor(Qr, device=dev), gidx < n_global / 2)
        del gidx
        n_units = n_global
        exchange_kind = {"nccl": "NCCL send/recv + all-reduce", "p2p-kernel": "one peer-memory kernel over NVLink (hsd_exchange_p2p)"}[sol.exchange] if world > 1 else "nothing (1 GPU)"
        local_cells = sol.nloc
        dx = 1.0 / n_global
        updated_local = sol.nloc - 2
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ------------------------------------------------------------
    for _ in range(args.warmup):
        sol.step(flux, 0.6, dx)
    torch.cuda.synchronize(); barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.15)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = kern.launches()
    torch.cuda.synchronize(); barrier()
    tw0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        sol.step(flux, 0.6, dx, kernel_events=ev[i])
    e1.record()
    torch.cuda.synchronize()
    tw1 = time.perf_counter()
    barrier()
    launches = kern.launches() - launches0
    ms_total = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], device=dev, dtype=torch.int64)
        dist.all_reduce(lt)
        launches = int(lt.item())
    clocks = sampler.stop(tw0, tw1) if sampler else None
    sol.check_status()
    value = n_units * args.steps / (ms_total * 1e-3)

    # ---- roofline of the fused step kernel -----------------------------------------------------
    peak, peak_src = measured_peaks()
    alg_bytes = 2 * nvar * 8 * updated_local
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic, fp64_pct = None, None
    tp = os.path.join(ROOT, "profiles", "step_kernel_traffic.json")
    if os.path.exists(tp):   # per-launch DRAM bytes / FP64-pipe utilisation of the same kernel from the committed ncu capture
        tj = json.load(open(tp)).get(args.workload)
        if tj and world == 1:
            traffic = tj.get("dram_bytes_per_launch")
            fp64_pct = tj.get("fp64_pipe_pct_of_peak")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": ("k_step_sp (TMA-fed tile pipeline: fused flux + update + next-step wave bounds)" if model == "sp13"
                           else "k_step (fused path-conservative flux + update + next-step wave bounds)"), "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_cell_update": 2 * nvar * 8, "cell_updates_per_launch": updated_local, "peak_source": peak_src,
                "fp64_pipe_pct_ncu": fp64_pct, "fp64_issue_peak_dfma_per_s": 1.708e13,
                "note": ("single-phase: DRAM traffic (algorithmic 208 B + 80 B of cached per-cell rows per cell update) and instruction issue are both near their limits (DESIGN.md)"
                         if model == "sp13" else "two-phase: FP64-pipe bound, not HBM bound (DESIGN.md): see profiles/ for the measured DFMA peak and pipe utilisation")}

    # ---- end to end through the host-buffer API ------------------------------------------------
    # every step: H2D of the whole (pinned) host state, CFL sweep, fused step, D2H of the new state.
    # Grid: the workload itself when it is <= 2^24 cells per GPU, else a 2^24-cell-per-GPU sample of it
    # (pinned host buffers of the 2^28 / ensemble workloads would not fit host memory twice).
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    del sol
    torch.cuda.empty_cache()
    if ensemble:
        nprob_e = min(wl["nprob"], (1 << 24) // wl["cells"] * world)
        pe0, pe1 = nprob_e * rank // world, nprob_e * (rank + 1) // world
        sol = EnsembleSolver(kern, wl["cells"], nprob_e)
        e_cells_local, e_units = sol.nprob * wl["cells"], nprob_e * wl["cells"]
        host_in = torch.empty(e_cells_local, nvar, dtype=torch.float64, pin_memory=True)
        host_out = torch.empty_like(host_in).pin_memory()
        left_h = (torch.arange(wl["cells"]) < wl["cells"] / 2)[None, :, None]
        q = torch.as_tensor(Qlr[: pe1 - pe0])
        host_in.view(sol.nprob, wl["cells"], nvar)[:] = torch.where(left_h, q[:, None, 0, :], q[:, None, 1, :])
        step_host = lambda a, b: sol.step_host(a, b, flux, 0.6, dx)
        e2e_api = "EnsembleSolver.step_host: pinned host state -> device, CFL sweep, fused step, device -> host, every step"
    else:
        e_cells = min(wl["cells"] if wl["scaling"] == "weak" else wl["cells"] // world, 1 << 24)
        e_units = e_cells * world
        if world == 1:
            s2 = H.Solver(eos, e_cells, model=hmodel, device=local)
            e_cells_local = e_cells
            lo_g = 0
            step_host = lambda a, b: s2.step_host(a.numpy(), b.numpy(), "hll", 0.6, 1.0 / e_units)
            e2e_api = "hs_step_host (C ABI): upload Q0 + CFL sweep + fused step + download Q1, every step"
        else:
            sol = SlabSolver(kern, e_units)
            e_cells_local, lo_g = sol.nloc, sol.lo_g
            step_host = lambda a, b: sol.step_host(a, b, flux, 0.6, 1.0 / e_units)
            e2e_api = "SlabSolver.step_host: pinned host slab -> device, CFL sweep + allreduce, fused step, halo, device -> host, every step"
        host_in = torch.empty(e_cells_local, nvar, dtype=torch.float64, pin_memory=True)
        host_out = torch.empty_like(host_in).pin_memory()
        gi = torch.arange(lo_g, lo_g + e_cells_local)
        host_in.copy_(torch.where((gi < e_units / 2)[:, None], torch.as_tensor(Ql)[None, :], torch.as_tensor(Qr)[None, :]))
    step_host(host_in, host_out)                     # warm-up (allocations, first touch)
    torch.cuda.synchronize(); barrier()
    l0 = kern.launches()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        step_host(host_in if i % 2 == 0 else host_out, host_out if i % 2 == 0 else host_in)
    torch.cuda.synchronize()
    e2e_t = time.perf_counter() - t0
    barrier()
    e2e_launches = kern.launches() - l0
    if world > 1:
        t = torch.tensor([e2e_t], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_t = float(t.item())
    nbytes_e2e = e_units * nvar * 8
    e2e = {"value": e_units * e2e_steps / e2e_t, "unit": UNIT, "h2d_bytes_per_step": nbytes_e2e, "d2h_bytes_per_step": nbytes_e2e,
           "steps": e2e_steps, "ms_per_step": 1e3 * e2e_t / e2e_steps, "cells_total": e_units, "api": e2e_api, "gpu_launches_per_rank": e2e_launches}

    # ---- CPU baseline (rank 0, N = 1 only) ------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as O
        threads = O.hardware_threads()
        v, sample, el = cpu_oracle_rate(model, args.cpu_seconds, threads, literal=True)
        v1, sample1, el1 = cpu_oracle_rate(model, min(4.0, args.cpu_seconds), 1, literal=True)   # also single-thread
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample, "seconds": el,
               "single_thread": {"value": v1, "sample": sample1, "seconds": el1},
               "note": "C++ restatement of main.jl (dual-number AD like ForwardDiff); Julia itself is not installed in this image"}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "model": None, "flux": "hll", "cfl": 0.6,
                       "cells_total": n_units, "cells_per_gpu": local_cells,
                       "parallelism": ("ensemble partition, no collective" if ensemble else f"slab x{world}, halo + max(lambda) per step via {exchange_kind}"),
                       "l2": f"state {local_cells * nvar * 8 / 1e9:.2f} GB per GPU per buffer >> 126 MB L2 (inputs larger than L2, no flush needed)"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        }
        out["config"].pop("model")
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sp13_2p24", choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
