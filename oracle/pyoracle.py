"""TEST INFRASTRUCTURE ONLY -- second, independent restatement of the reference's constitutive
chain, used to pin the C++ oracle (oracle.cpp) and to generate tests/golden/*.json.

Independence: derivatives come from torch *reverse-mode* autograd (the C++ oracle uses
forward-mode duals, the CUDA kernels closed forms), determinants from torch.linalg.det (LAPACK
LU), eigenvalues from numpy.linalg.eigvals (LAPACK dgeev on the un-symmetrised matrix, the same
routine Julia's `eigvals` dispatches to for a general real matrix).

Each function cites the reference file:line it follows.  PARITY UNPINNED against the real
Julia reference (no Julia here, no golden vectors shipped) -- see oracle/README.md.

Run `python oracle/pyoracle.py` to (re)generate tests/golden/pyoracle_vectors.json and pyoracle_hank_vectors.json.
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

torch.set_default_dtype(torch.float64)


class Barton2009:
    """EquationsOfState.jl:71-116"""

    def __init__(self, rho0=8.93, c0=4.6, cv=3.9e-4, t0=300, b0=2.1, alpha=1, beta=3, gamma=2):
        self.rho0, self.c0, self.cv, self.t0, self.b0 = rho0, c0, cv, t0, b0
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.b0sq = b0 ** 2
        self.k0 = c0 ** 2 - (4 / 3) * b0 ** 2

    def block(self):
        return [float(v) for v in (self.rho0, self.c0, self.cv, self.t0, self.b0, self.alpha, self.beta, self.gamma, self.b0sq, self.k0)]


def mat(v):  # reshape(v, (3,3)) of a column-major 9-vector
    return v.reshape(3, 3).T


def vec(m):  # m[:]
    return m.T.reshape(9)


def finger(a):  # Strains.jl:26-32
    A = mat(a)
    return vec(torch.linalg.inv(A @ A.T))


def invariants(g):  # Strains.jl:46-52
    G = mat(g)
    i1 = torch.trace(G)
    i2 = 0.5 * (torch.trace(G) ** 2 - torch.trace(G @ G))
    i3 = torch.linalg.det(G)
    return i1, i2, i3


def energy(eos, S, G):  # EquationsOfState.jl:118-137
    i1, i2, i3 = invariants(G)
    U = (0.5 * eos.k0 / (eos.alpha ** 2) * (i3 ** (0.5 * eos.alpha) - 1.0) ** 2
         + eos.cv * eos.t0 * i3 ** (0.5 * eos.gamma) * (torch.exp(S / eos.cv) - 1.0))
    W = 0.5 * eos.b0sq * i3 ** (0.5 * eos.beta) * (i1 ** 2 / 3.0 - i2)
    return U + W


def entropy(eos, e_int, G):  # EquationsOfState.jl:139-156
    i1, i2, i3 = invariants(G)
    S = e_int - 0.5 * eos.b0sq * i3 ** (0.5 * eos.beta) * (i1 ** 2 / 3 - i2) - 0.5 * eos.k0 / (eos.alpha ** 2) * (i3 ** (0.5 * eos.alpha) - 1) ** 2
    S = S / (eos.cv * eos.t0 * i3 ** (0.5 * eos.gamma)) + 1
    if S < 1e-6:
        S = torch.tensor(1e-6)
    return torch.log(S) * eos.cv


def stress(eos, ent, F, create_graph=False):  # EquationsOfState.jl:179-190
    den = eos.rho0 / torch.linalg.det(mat(F))
    G = finger(F)
    if not G.requires_grad:
        G = G.detach().requires_grad_(True)
        Gv = G
    else:
        Gv = G
    e = energy(eos, ent, Gv)
    (dedG,) = torch.autograd.grad(e, Gv, create_graph=create_graph)
    return vec(-2 * den * (mat(G) @ mat(dedG)))


def acoustic(eos, ent, F, n=(1.0, 0.0, 0.0)):  # EquationsOfState.jl:223-246
    Fv = F.detach().clone().requires_grad_(True)

    def f(x):
        den = eos.rho0 / torch.linalg.det(mat(x))
        G = finger(x)
        e = energy(eos, ent, G)
        (dedG,) = torch.autograd.grad(e, G, create_graph=True)
        return vec(-2 * den * (mat(G) @ mat(dedG)))

    J = torch.autograd.functional.jacobian(f, Fv)  # [sigma idx (m + 3 i), F idx (j + 3 l)]
    dTdF = J.numpy().reshape(3, 3, 3, 3, order="F")  # reshape(jacobian, (3,3,3,3)) -> [m, i, j, l]
    Fm = mat(F).numpy()
    den = eos.rho0 / np.linalg.det(Fm)
    A = (1 / den) * dTdF
    ac = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    for m in range(3):
                        ac[i, j] += A[m, i, j, l] * Fm[k, l] * n[m] * n[k]
    return ac


def cons2prim(eos, Q):  # HyperelasticityMPh.jl:106-133
    frac = Q[0]
    FQ = mat(Q[6:15] / frac)
    true_den = torch.sqrt(torch.linalg.det(FQ) / eos.rho0)
    den = frac * true_den
    vel = Q[2:5] / den
    e_total = Q[5] / den
    e_int = e_total - (vel ** 2).sum() / 2
    F = Q[6:15] / den
    ent = entropy(eos, e_int, finger(F))
    return torch.cat([torch.stack([frac, true_den]), vel, ent.reshape(1), F])


def prim2cons(eos, P):  # HyperelasticityMPh.jl:66-87
    frac, true_den = P[0], P[1]
    den = frac * true_den
    vel, ent, F = P[2:5], P[5], P[6:15]
    e_total = energy(eos, ent, finger(F)) + (vel ** 2).sum() / 2
    return torch.cat([torch.stack([frac, den]), den * vel, (den * e_total).reshape(1), den * F])


def flux(eos, Q):  # HyperelasticityMPh.jl:146-175
    P = cons2prim(eos, Q)
    frac, true_den, vel, ent, F = P[0], P[1], P[2:5], P[5], P[6:15]
    den = frac * true_den
    e_total = Q[5] / den
    strs = frac * stress(eos, ent.detach(), F.detach())
    f = torch.zeros(15)
    f[1] = den * vel[0]
    f[2:5] = den * vel[0] * vel - strs[0::3]
    f[5] = den * vel[0] * e_total - (vel * strs[0::3]).sum()
    f[6:15] = den * (vel[0] * F - vec(torch.outer(vel, F[0::3])))
    return f


def get_eigvals(eos, Q):  # HyperelasticityMPh.jl:258-266, n = (1,0,0)
    P = cons2prim(eos, Q)
    ac = acoustic(eos, P[5].detach(), P[6:15].detach())
    ev = np.linalg.eigvals(ac)                      # LAPACK dgeev, un-symmetrised
    c = np.sort(np.sqrt(np.abs(ev)))
    spd = float(P[2])
    return np.concatenate([spd + c, spd - c]), ac


def noncons_cols(eoss, Q):  # HyperelasticityMPh.jl:178-250 (column 1 of each block)
    ph = []
    for p in range(2):
        q = Q[15 * p:15 * p + 15]
        P = cons2prim(eoss[p], q)
        frac, true_den, vel, ent, F = P[0], P[1], P[2:5], P[5].detach(), P[6:15].detach()
        strs = mat(frac * stress(eoss[p], ent, F))
        S = ent.clone().requires_grad_(True)
        (temp,) = torch.autograd.grad(energy(eoss[p], S, finger(F)), S)   # :212
        ph.append(dict(frac=frac, true_den=true_den, vel=vel, F=F, strs=strs, temp=temp))
    k = [0.5, 0.5]
    vel_i = k[0] * ph[0]["vel"] + k[1] * ph[1]["vel"]
    K = [1 / ph[p]["frac"] * ph[p]["strs"] for p in range(2)]                      # :216 with omega = 0
    strs_i = (k[1] * ph[1]["temp"] * K[0] + k[0] * ph[0]["temp"] * K[1]) / (k[0] * ph[0]["temp"] + k[1] * ph[1]["temp"])
    cols = []
    for p in range(2):
        c = torch.zeros(15)
        c[0] = vel_i[0]
        c[2:5] = strs_i[:, 0]
        c[5] = (strs_i[:, 0] * vel_i).sum()
        F, vel, rho = ph[p]["F"], ph[p]["vel"], ph[p]["true_den"]
        for i in (0, 3, 6):
            c[6 + i:9 + i] = rho * F[i] * vel
        c[6:15:3] += rho * (mat(F).T @ (vel_i - vel))
        cols.append(c)
    return torch.cat(cols)


# ---------------------------------------------------------------------------------------------
# Hank2016, EquationsOfState.jl:301-356 (dead code in the reference; a 3x3 tensor is read as its 9 column-major
# entries, see oracle.cpp).  Gradient by reverse-mode autograd.
class Hank2016:
    def __init__(self, rho0=2.7, mu=26e9, gamma=3.4, pres_inf=21.5e9, a=0.5):
        self.rho0, self.mu, self.gamma, self.pres_inf, self.a = rho0, mu, gamma, pres_inf, a

    def block(self):
        return [float(v) for v in (self.rho0, self.mu, self.gamma, self.pres_inf, self.a)]


def hank_e_el(eos, i1, i2, i3):  # EquationsOfState.jl:326-328
    j1 = i1 / i3 ** (1 / 3)
    j2 = (i1 ** 2 - 2 * i2) / i3 ** (2 / 3)
    return eos.mu / (4 * eos.rho0) * ((1 - 2 * eos.a) / 3 * j1 ** 2 + eos.a * j2 + 3 * (eos.a - 1))


def hank_energy(eos, den, pres, G):  # EquationsOfState.jl:317-331
    return hank_e_el(eos, *invariants(G)) + (pres + eos.gamma * eos.pres_inf) / (den * (eos.gamma - 1))


def hank_pressure(eos, den, e_int, inv3):  # EquationsOfState.jl:333-346
    return (e_int - hank_e_el(eos, *inv3)) * (eos.gamma - 1) * den - eos.gamma * eos.pres_inf


def hank_stress(eos, den, pres, distortion):  # EquationsOfState.jl:348-356
    G = finger(vec(torch.linalg.inv(mat(distortion)))).detach().requires_grad_(True)
    e = hank_energy(eos, den, pres, G)
    (dedG,) = torch.autograd.grad(e, G)
    return vec(-2.0 * den * mat(G.detach()) @ mat(dedG))


def generate_hank(path):
    rng = np.random.default_rng(20261018)
    doc = {"generator": "oracle/pyoracle.py::generate_hank (torch reverse-mode autograd)", "cases": []}
    for eos in (Hank2016(), Hank2016(rho0=8.9, mu=48e9, gamma=4.2, pres_inf=34e9, a=-0.3)):
        for k in range(6):
            A = np.eye(3) + 0.2 * rng.uniform(-1, 1, (3, 3))
            a9 = torch.tensor(A.flatten(order="F"))
            den = float(eos.rho0 * abs(np.linalg.det(A)) * rng.uniform(0.9, 1.1)); pres = float(rng.uniform(-1e9, 5e10))
            g9 = vec(mat(a9).T @ mat(a9))
            i1, i2, i3 = invariants(g9)
            e = hank_energy(eos, den, pres, g9)
            doc["cases"].append({"eos_block": eos.block(), "distortion": a9.tolist(), "den": den, "pres": pres, "G": g9.tolist(),
                                 "invariants": [float(i1), float(i2), float(i3)], "energy": float(e),
                                 "pressure": float(hank_pressure(eos, den, e, (i1, i2, i3))),
                                 "stress": hank_stress(eos, den, pres, a9).tolist()})
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    return doc


# ---------------------------------------------------------------------------------------------
def _states():
    rng = np.random.default_rng(20261017)
    out = []
    for k in range(6):
        a1 = rng.uniform(0.1, 0.9)
        P = []
        u = rng.uniform(-1, 1, 3); S = rng.uniform(0, 1e-3)
        F = np.eye(3) + 0.05 * rng.uniform(-1, 1, (3, 3))
        for a in (a1, 1 - a1):
            P += [a, 8.9 / np.linalg.det(F), *u, S, *F.flatten(order="F")]
        out.append(P)
    return np.array(out)


def generate(path):
    eos_sets = {
        "default": (Barton2009(), Barton2009()),
        "hetero": (Barton2009(), Barton2009(rho0=8.93, c0=6.22, cv=9.0e-4, t0=300, b0=3.16, alpha=1, beta=3.577, gamma=2.088)),
    }
    doc = {"generator": "oracle/pyoracle.py (torch reverse-mode autograd + numpy.linalg.eigvals)", "cases": []}
    for name, eoss in eos_sets.items():
        for P in _states():
            Pt = torch.tensor(P)
            Q = torch.cat([prim2cons(eoss[p], Pt[15 * p:15 * p + 15]) for p in range(2)]).detach()
            case = {"eos": name, "eos_blocks": [e.block() for e in eoss], "P": P.tolist(), "Q": Q.tolist()}
            case["cons2prim"] = torch.cat([cons2prim(eoss[p], Q[15 * p:15 * p + 15]) for p in range(2)]).detach().tolist()
            case["flux"] = torch.cat([flux(eoss[p], Q[15 * p:15 * p + 15]) for p in range(2)]).detach().tolist()
            eg, acs = [], []
            for p in range(2):
                e, ac = get_eigvals(eoss[p], Q[15 * p:15 * p + 15])
                eg += e.tolist(); acs.append(ac.tolist())
            case["eigvals"] = eg
            case["acoustic"] = acs
            case["noncons_cols"] = noncons_cols(eoss, Q).detach().tolist()
            doc["cases"].append(case)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    return doc


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(os.path.dirname(here), "tests", "golden", "pyoracle_vectors.json")
    d = generate(p)
    print("wrote", p, len(d["cases"]), "cases")
    p = os.path.join(os.path.dirname(here), "tests", "golden", "pyoracle_hank_vectors.json")
    d = generate_hank(p)
    print("wrote", p, len(d["cases"]), "cases")
