"""Host-side driver pieces of main.jl that surround the hot path:
CSV snapshots compatible with the reference's `save_data` / `read_data` / restart logic and with
`plotter.py`, the tanh-smoothed initial condition, and the `while t < T` driver itself.

    python -m hyperelasticsolver_b200.driver --testcase 6 --nx 1000 --T 0.06

reproduces the shipped default run (main.jl:133-152) with the state resident on the GPU between
snapshots.  Everything numerical goes through the C ABI; this module only moves text and files.
"""
from __future__ import annotations

import argparse
import logging
import math
import os

import numpy as np

from . import _lib as L
from .hyperelasticity_mph import cons2prim_mph, initial_states, prim2cons_mph
from .solver import Solver, initial_condition

# main.jl:71
HEADER = ("a1\tr1\tu11\tu21\tu31\tS1\tF111\tF211\tF311\tF121\tF221\tF321\tF131\tF231\tF331\t"
          "a2\tr2\tu12\tu22\tu32\tS2\tF112\tF212\tF312\tF122\tF222\tF322\tF132\tF232\tF332")


def julia_float(x: float) -> str:
    """Text of a Float64 as Julia's `print` writes it (shortest round-trip digits; exponent form
    below 1e-5 and from 1e6 up; `1.0e-5`, not `1e-05`), so files are byte-compatible with
    `join(P, "\\t")` of main.jl:74."""
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    r = repr(float(x))
    sign = "-" if r.startswith("-") else ""
    r = r.lstrip("-")
    if "e" in r:
        mant, exp = r.split("e")
        digits = mant.replace(".", "")
        e10 = int(exp) + (len(mant.split(".")[0]) - 1)
    else:
        ip, _, fp = r.partition(".")
        fp = fp if fp != "0" else ""
        if ip.strip("0"):
            digits = (ip.lstrip("0") + fp).rstrip("0") or "0"
            e10 = len(ip.lstrip("0")) - 1
        else:
            nz = len(fp) - len(fp.lstrip("0"))
            digits = fp.lstrip("0").rstrip("0")
            e10 = -nz - 1
    digits = digits.rstrip("0") or "0"
    if -5 < e10 < 6:
        if e10 >= 0:
            ip = digits[:e10 + 1].ljust(e10 + 1, "0")
            fp = digits[e10 + 1:] or "0"
            return f"{sign}{ip}.{fp}"
        return f"{sign}0.{'0' * (-e10 - 1)}{digits}"
    mant = digits[0] + "." + (digits[1:] or "0")
    return f"{sign}{mant}e{e10}"


def get_filename(step_num: int) -> str:
    """main.jl:108"""
    return "sol_%06i.csv" % step_num


def save_data(fname: str, Q, eos, device=0):
    """main.jl:67-77: primitives of every cell, tab separated, under the fixed 30-name header."""
    P = cons2prim_mph(eos, np.ascontiguousarray(Q, dtype=np.float64), device=device)
    with open(fname, "w") as io:
        io.write(HEADER + "\n")
        for row in P:
            io.write("\t".join(julia_float(float(v)) for v in row) + "\n")


def read_data(fname: str):
    """main.jl:84-92 -> (P (nx, 30), nx)."""
    with open(fname) as f:
        lines = f.read().splitlines()
    P = np.array([[float(x) for x in ln.split()] for ln in lines[1:] if ln.strip()], dtype=np.float64)
    return P, P.shape[0]


def initial_condition_tanh(eos, Ql, Qr, nx, eps, device=0):
    """main.jl:110-123: two states with a tanh-smoothed volume-fraction profile."""
    Pl = cons2prim_mph(eos, Ql, device=device)
    Pr = cons2prim_mph(eos, Qr, device=device)
    x = (np.arange(1, nx + 1) - 0.5) / nx
    P = np.where((x < 0.5)[:, None], Pl[None, :], Pr[None, :]).copy()
    P[:, 0] = 0.2 / 2 * (np.tanh(4 * (x - 0.5) / eps) + 1) + 0.4
    P[:, 15] = 1 - P[:, 0]
    return prim2cons_mph(eos, P, device=device)


def run(eos=None, testcase=6, nx=1000, cfl=0.6, T=0.06, X=1.0, log_freq=100, dir_name="barton_data/", flux="hll",
        device=0, tanh_eps=None, log=None):
    """The script body of main.jl:133-245 (logging to `solution.log` is left to the caller's logger)."""
    log = log or logging.getLogger("hyperelasticsolver_b200")
    eos = eos or (L.Barton2009(), L.Barton2009())
    dx = X / nx
    dt_const = 5 * 1e-6                       # main.jl:145 (only used by the restart clock, as in the reference)
    t, step_num = 0.0, 0
    os.makedirs(dir_name, exist_ok=True)
    log.info("Data directory: %s", dir_name)
    files = sorted(os.listdir(dir_name))
    if "result.csv" in files or not files:    # main.jl:174-180
        for f in files:
            os.remove(os.path.join(dir_name, f))
        log.info("Cleaning data directory: %s", dir_name)
        Ql, Qr = initial_states(eos, testcase, device=device)
        Q0 = initial_condition(Ql, Qr, nx) if tanh_eps is None else initial_condition_tanh(eos, Ql, Qr, nx, tanh_eps, device)
    else:                                     # main.jl:181-193
        last_file = os.path.join(dir_name, files[-1])
        log.info("Found file: %s", last_file)
        step_num = int(os.path.basename(last_file).split(".")[0].split("_")[1])
        t = step_num * dt_const
        P0, nx = read_data(last_file)
        Q0 = prim2cons_mph(eos, P0, device=device)
    fname = os.path.join(dir_name, get_filename(step_num))
    save_data(fname, Q0, eos, device)
    log.info("Initial state saved to: %s", fname)

    with Solver(eos, nx, model=L.MPH30, device=device) as sol:
        sol.upload(Q0)
        sol.set_time(t, step_num)
        while sol.t[0] < T:                   # main.jl:202
            # stay on the device until the next snapshot (main.jl:233: step_num % log_freq == 0)
            todo = log_freq - int(sol.steps[0]) % log_freq
            hist = sol.advance(T, flux, cfl, dx, max_steps=todo, record_dt=True)
            taken = int((hist[0] > 0).sum())
            tt = sol.t[0] - hist[0, :taken].sum()
            for k in range(taken):
                tt += hist[0, k]
                log.info("Step = %d,\t t = %.6f / %.6f,\t Δt = %.6f", int(sol.steps[0]) - taken + k + 1, tt, T, hist[0, k])
            if int(sol.steps[0]) % log_freq == 0 and taken:
                fname = os.path.join(dir_name, get_filename(int(sol.steps[0])))
                save_data(fname, sol.download(), eos, device)
                log.info("Solution saved to: %s", fname)
        Q = sol.download()
        t, step_num = float(sol.t[0]), int(sol.steps[0])
    fname = os.path.join(dir_name, "result.csv")
    save_data(fname, Q, eos, device)
    log.info("Result solution saved to: %s", fname)
    log.info("Done!")
    return Q, t, step_num


def main(argv=None):
    ap = argparse.ArgumentParser(description="GPU drop-in for `julia main.jl` (two-phase model, main.jl:133-152 defaults)")
    ap.add_argument("--testcase", type=int, default=6)
    ap.add_argument("--nx", type=int, default=1000)
    ap.add_argument("--cfl", type=float, default=0.6)
    ap.add_argument("--T", type=float, default=0.06)
    ap.add_argument("--flux", default="hll", choices=["hll", "lxf"])
    ap.add_argument("--dir", default="barton_data/")
    ap.add_argument("--log-freq", type=int, default=100)
    ap.add_argument("--tanh-eps", type=float, default=None, help="use initial_condition_tanh (main.jl:180) with this eps")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(message)s",
                        handlers=[logging.StreamHandler(), logging.FileHandler("solution.log")])   # main.jl:157-163
    run(None, a.testcase, a.nx, a.cfl, a.T, 1.0, a.log_freq, a.dir, a.flux, a.device, a.tanh_eps)


if __name__ == "__main__":
    main()
