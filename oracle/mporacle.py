"""TEST INFRASTRUCTURE ONLY -- third restatement of the reference's constitutive chain, in ARBITRARY PRECISION
(mpmath, 50 digits), to bound the FP64 roundoff of the C++ oracle itself.

The reference ships no golden vectors and Julia is not installed, so parity is "unpinned" against the real program
(oracle/README.md).  What can be quantified is the other half of the argument "every quantity on the path has a
mathematically unique value, so any two correct implementations agree to roundoff": this file evaluates the same
formulas -- literally, function for function, with forward-mode dual numbers where the reference uses ForwardDiff
(nested for `acoustic`) -- in 50-digit arithmetic, where roundoff is ~1e-50.  tests/test_oracle_golden.py asserts that
the C++ oracle (FP64) is within a few 1e-15 of these values on the 24 golden states and on one path-conservative HLL face.
An FP64 implementation that agrees with the oracle to 1e-12 therefore agrees with EXACT arithmetic on the reference's
formulas to 1e-12 as well; what remains unpinned is only whether the formulas were read correctly (two people, three
restatements: C++ duals, torch reverse mode, this one).

Each function cites the reference file:line it follows.  Nothing under hyperelasticsolver_b200/ imports this.

  python oracle/mporacle.py      (re)generates tests/golden/mp_vectors.json
"""
from __future__ import annotations

import json
import os

import mpmath as mp

mp.mp.dps = 50
mpf = mp.mpf


# ---------------------------------------------------------------------------------------------
# forward-mode dual numbers over any field (mpf, or Dual again for second derivatives): ForwardDiff's Dual
class Dual:
    __slots__ = ("v", "d")

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def lift(x, like):
        return x if isinstance(x, Dual) else Dual(x, like.d * 0)

    def __add__(self, o):
        o = Dual.lift(o, self); return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__

    def __neg__(self):
        return Dual(-self.v, -self.d)

    def __sub__(self, o):
        o = Dual.lift(o, self); return Dual(self.v - o.v, self.d - o.d)

    def __rsub__(self, o):
        return Dual.lift(o, self) - self

    def __mul__(self, o):
        o = Dual.lift(o, self); return Dual(self.v * o.v, self.v * o.d + self.d * o.v)
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o, self); return Dual(self.v / o.v, (self.d * o.v - self.v * o.d) / (o.v * o.v))

    def __rtruediv__(self, o):
        return Dual.lift(o, self) / self

    def __pow__(self, p):   # real exponent
        return Dual(power(self.v, p), p * power(self.v, p - 1) * self.d)

    def __lt__(self, o):
        return value(self) < value(o)


def value(x):
    while isinstance(x, Dual):
        x = x.v
    return x


def power(x, p):
    return x ** p if isinstance(x, Dual) else mp.power(x, p)


def exp(x):
    if isinstance(x, Dual):
        e = exp(x.v); return Dual(e, e * x.d)
    return mp.exp(x)


def log(x):
    if isinstance(x, Dual):
        return Dual(log(x.v), x.d / x.v)
    return mp.log(x)


def sqrt(x):
    if isinstance(x, Dual):
        s = sqrt(x.v); return Dual(s, x.d / (2 * s))
    return mp.sqrt(x)


# ---------------------------------------------------------------------------------------------
# 3x3 algebra on lists of 9 (column-major, v[i + 3 j] = M[i][j]) -- SimpleLA.jl:51-96
def mat(v):
    return [[v[i + 3 * j] for j in range(3)] for i in range(3)]


def vec(M):
    return [M[i][j] for j in range(3) for i in range(3)]


def matmul(A, B):
    return [[A[i][0] * B[0][j] + A[i][1] * B[1][j] + A[i][2] * B[2][j] for j in range(3)] for i in range(3)]


def transpose(A):
    return [[A[j][i] for j in range(3)] for i in range(3)]


def det3(A):   # det_, SimpleLA.jl:84-89
    return (A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0])
            + A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]))


def inv3(A):   # inv_, SimpleLA.jl:51-79 (cofactors / det)
    d = det3(A)
    C = [[None] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            r = [k for k in range(3) if k != i]; c = [k for k in range(3) if k != j]
            m = A[r[0]][c[0]] * A[r[1]][c[1]] - A[r[0]][c[1]] * A[r[1]][c[0]]
            C[j][i] = (m if (i + j) % 2 == 0 else -m) / d
    return C


def finger(a):   # Strains.jl:26-32
    A = mat(a)
    return vec(inv3(matmul(A, transpose(A))))


def invariants(g):   # Strains.jl:46-52
    G = mat(g)
    tr = G[0][0] + G[1][1] + G[2][2]
    G2 = matmul(G, G)
    tr2 = G2[0][0] + G2[1][1] + G2[2][2]
    return tr, (tr * tr - tr2) / 2, det3(G)


class Barton2009:   # EquationsOfState.jl:71-116
    def __init__(self, block):
        (self.rho0, self.c0, self.cv, self.t0, self.b0, self.alpha, self.beta, self.gamma, self.b0sq, self.k0) = [mpf(float(x)) for x in block]
        # the derived fields are recomputed exactly from the primary ones (EquationsOfState.jl:111-112): the FP64 struct
        # carries b0^2 and c0^2 - 4/3 b0^2 rounded; the rounded values ARE the reference's parameters, keep them
        self.half = mpf(1) / 2


def energy(eos, S, G):   # EquationsOfState.jl:118-137
    i1, i2, i3 = invariants(G)
    U = (eos.k0 / (2 * eos.alpha ** 2) * (power(i3, eos.alpha / 2) - 1) ** 2
         + eos.cv * eos.t0 * power(i3, eos.gamma / 2) * (exp(S / eos.cv) - 1))
    W = eos.b0sq / 2 * power(i3, eos.beta / 2) * (i1 ** 2 / 3 - i2)
    return U + W


def entropy(eos, e_int, G):   # EquationsOfState.jl:139-156 (clamp :152-154)
    i1, i2, i3 = invariants(G)
    S = e_int - eos.b0sq / 2 * power(i3, eos.beta / 2) * (i1 ** 2 / 3 - i2) - eos.k0 / (2 * eos.alpha ** 2) * (power(i3, eos.alpha / 2) - 1) ** 2
    S = S / (eos.cv * eos.t0 * power(i3, eos.gamma / 2)) + 1
    if value(S) < mpf(1e-6):          # (the reference's literal is the double nearest to 1e-6)
        S = mpf(1e-6) + S * 0
    return log(S) * eos.cv


def gradient9(f, x):   # ForwardDiff.gradient over a 9-vector
    out = []
    zero = x[0] * 0
    for k in range(9):
        xd = [Dual(x[i], (zero + 1) if i == k else zero) for i in range(9)]
        out.append(f(xd).d)
    return out


def stress(eos, ent, F):   # EquationsOfState.jl:179-190
    den = eos.rho0 / det3(mat(F))
    G = finger(F)
    dedG = gradient9(lambda g: energy(eos, ent, g), G)
    M = matmul(mat(G), mat(dedG))
    return [-2 * den * x for x in vec(M)]


def acoustic(eos, ent, F, n=(1, 0, 0)):   # EquationsOfState.jl:223-246: jacobian of stress w.r.t. F (nested duals), then the contraction
    J = [[None] * 9 for _ in range(9)]   # J[sigma idx][F idx]
    for k in range(9):
        Fd = [Dual(F[i], mpf(1) if i == k else mpf(0)) for i in range(9)]
        sd = stress(eos, ent, Fd)
        for r in range(9):
            J[r][k] = sd[r].d
    Fm = mat(F)
    den = eos.rho0 / det3(Fm)
    ac = [[mpf(0)] * 3 for _ in range(3)]
    for i in range(3):
        for j in range(3):
            for k in range(3):
                for l in range(3):
                    for m in range(3):
                        # reshape(jacobian, (3,3,3,3))[m, i, j, l] = J[m + 3 i][j + 3 l]
                        ac[i][j] += (1 / den) * J[m + 3 * i][j + 3 * l] * Fm[k][l] * n[m] * n[k]
    return ac


def cons2prim(eos, Q):   # HyperelasticityMPh.jl:106-133
    frac = Q[0]
    FQ = mat([x / frac for x in Q[6:15]])
    true_den = sqrt(det3(FQ) / eos.rho0)
    den = frac * true_den
    vel = [x / den for x in Q[2:5]]
    e_total = Q[5] / den
    e_int = e_total - (vel[0] ** 2 + vel[1] ** 2 + vel[2] ** 2) / 2
    F = [x / den for x in Q[6:15]]
    ent = entropy(eos, e_int, finger(F))
    return [frac, true_den] + vel + [ent] + F


def flux(eos, Q):   # HyperelasticityMPh.jl:146-175
    P = cons2prim(eos, Q)
    frac, true_den, vel, ent, F = P[0], P[1], P[2:5], P[5], P[6:15]
    den = frac * true_den
    e_total = Q[5] / den
    strs = [frac * x for x in stress(eos, ent, F)]
    row1 = strs[0::3]
    f = [mpf(0)] * 15
    f[1] = den * vel[0]
    for k in range(3):
        f[2 + k] = den * vel[0] * vel[k] - row1[k]
    f[5] = den * vel[0] * e_total - (vel[0] * row1[0] + vel[1] * row1[1] + vel[2] * row1[2])
    F1 = F[0::3]   # row 1 of F
    for j in range(3):
        for i in range(3):
            f[6 + i + 3 * j] = den * (vel[0] * F[i + 3 * j] - vel[i] * F1[j])
    return f


def eig_sym3(ac):   # eigvals of the (symmetric in exact arithmetic) acoustic tensor, HyperelasticityMPh.jl:263
    A = mp.matrix([[(ac[i][j] + ac[j][i]) / 2 for j in range(3)] for i in range(3)])
    E = mp.eigsy(A, eigvals_only=True)
    return sorted([E[k] for k in range(3)])


def get_eigvals_phase(eos, Q):   # HyperelasticityMPh.jl:258-266, n = (1,0,0)
    P = cons2prim(eos, Q)
    ac = acoustic(eos, P[5], P[6:15])
    c = sorted(mp.sqrt(abs(e)) for e in eig_sym3(ac))
    return [P[2] + x for x in c] + [P[2] - x for x in c], ac


def noncons_cols(eoss, Q):   # HyperelasticityMPh.jl:178-250 (column 1 of each 15x15 block; omega = 0, k = (1/2, 1/2))
    ph = []
    for p in range(2):
        q = Q[15 * p:15 * p + 15]
        P = cons2prim(eoss[p], q)
        frac, true_den, vel, ent, F = P[0], P[1], P[2:5], P[5], P[6:15]
        strs = mat([frac * x for x in stress(eoss[p], ent, F)])
        temp = energy(eoss[p], Dual(ent, mpf(1)), finger(F)).d     # ForwardDiff.derivative(S -> energy(eos, S, G)), :212
        ph.append((frac, true_den, vel, F, strs, temp))
    k = [mpf(1) / 2, mpf(1) / 2]
    vel_i = [k[0] * ph[0][2][i] + k[1] * ph[1][2][i] for i in range(3)]
    K = [[[ph[p][4][i][j] / ph[p][0] for j in range(3)] for i in range(3)] for p in range(2)]
    T0, T1 = ph[0][5], ph[1][5]
    strs_i = [[(k[1] * T1 * K[0][i][j] + k[0] * T0 * K[1][i][j]) / (k[0] * T0 + k[1] * T1) for j in range(3)] for i in range(3)]
    cols = []
    for p in range(2):
        frac, rho, vel, F, _, _ = ph[p]
        c = [mpf(0)] * 15
        c[0] = vel_i[0]
        for i in range(3):
            c[2 + i] = strs_i[i][0]
        c[5] = strs_i[0][0] * vel_i[0] + strs_i[1][0] * vel_i[1] + strs_i[2][0] * vel_i[2]
        for blk in (0, 3, 6):
            for kk in range(3):
                c[6 + blk + kk] = rho * F[blk] * vel[kk]
        Fm = mat(F)
        dv = [vel_i[i] - vel[i] for i in range(3)]
        for j in range(3):
            c[6 + 3 * j] += rho * (Fm[0][j] * dv[0] + Fm[1][j] * dv[1] + Fm[2][j] * dv[2])
        cols += c
    return cols


def gauss_legendre6():   # gausslegendre(6) mapped to [0,1], NumFluxes.jl:94-95
    x = mp.polyroots([231, 0, -315, 0, 105, 0, -5], maxsteps=200, extraprec=200)   # 16 P6(x) = 231 x^6 - 315 x^4 + 105 x^2 - 5
    x = sorted(mp.re(r) for r in x)
    dP = lambda t: (1386 * t ** 5 - 1260 * t ** 3 + 210 * t) / 16
    w = [2 / ((1 - t * t) * dP(t) ** 2) for t in x]
    return [(t + 1) / 2 for t in x], [wi / 2 for wi in w]


def hll_pathcons(eoss, Ql, Qr, eig_l, eig_r):   # NumFluxes.jl:85-132
    def speeds(Q):
        out = []
        for p in range(2):
            e, _ = get_eigvals_phase(eoss[p], Q[15 * p:15 * p + 15])
            out += e
        return out

    def flux_mph(Q):
        return flux(eoss[0], Q[:15]) + flux(eoss[1], Q[15:])

    Qm = [(a + b) / 2 for a, b in zip(Ql, Qr)]
    em = speeds(Qm)
    s_l = min([mpf(0), min(em), min(eig_l)])
    s_r = max([mpf(0), max(em), max(eig_r)])
    xs, ws = gauss_legendre6()

    def B_int(Qa, Qb):
        acc = [mpf(0)] * 30
        for x, w in zip(xs, ws):
            psi = [a * (1 - x) + b * x for a, b in zip(Qa, Qb)]
            cols = noncons_cols(eoss, psi)
            for p in range(2):
                dalpha = Qb[15 * p] - Qa[15 * p]
                for r in range(15):
                    acc[15 * p + r] += w * cols[15 * p + r] * dalpha
        return acc

    Fl, Fr = flux_mph(Ql), flux_mph(Qr)
    B1 = B_int(Ql, Qr)
    path = [B1[i] + Fr[i] - Fl[i] for i in range(30)]
    Qh = [(Qr[i] * s_r - Ql[i] * s_l - path[i]) / (s_r - s_l) for i in range(30)]
    B2, B3 = B_int(Ql, Qh), B_int(Qh, Qr)
    dm = [-s_l / (s_r - s_l) * (Fr[i] - Fl[i] + B2[i] + B3[i]) + s_l * s_r / (s_r - s_l) * (Qr[i] - Ql[i]) for i in range(30)]
    dp = [s_r / (s_r - s_l) * (Fr[i] - Fl[i] + B2[i] + B3[i]) - s_l * s_r / (s_r - s_l) * (Qr[i] - Ql[i]) for i in range(30)]
    return s_l, s_r, dm, dp


def time_step(eoss, cells, cfl, dx):
    """one pass of the loop body of main.jl:204-227 on a list of 30-vectors: CFL sweep (:204-212), frozen boundary cells (:219-220),
    update_cell with hll (:43-60, :225).  Returns (new cells, dt)."""
    eig = []
    for Q in cells:
        e = []
        for p in range(2):
            e += get_eigvals_phase(eoss[p], Q[15 * p:15 * p + 15])[0]
        eig.append(e)
    lam = max(max(abs(x) for x in e) for e in eig)               # main.jl:210-212
    dt = cfl * dx / lam
    dtdx = dt / dx
    faces = [hll_pathcons(eoss, cells[i], cells[i + 1], eig[i], eig[i + 1]) for i in range(len(cells) - 1)]   # (s_l, s_r, D-, D+)
    new = [list(cells[0])]
    for i in range(1, len(cells) - 1):
        NF_l = faces[i - 1][3]      # D+ of the left face   (main.jl:56)
        NF_r = faces[i][2]          # D- of the right face  (main.jl:57)
        new.append([cells[i][k] - dtdx * ((0 - 0) + (NF_r[k] + NF_l[k])) for k in range(30)])        # main.jl:59 (hll's conservative part is zero)
    new.append(list(cells[-1]))
    return new, dt


# ---------------------------------------------------------------------------------------------
def s30(x):
    return mp.nstr(x, 30, strip_zeros=False, min_fixed=-1000, max_fixed=-999)   # 30 significant digits, exponent form


def generate(path):
    here = os.path.dirname(os.path.abspath(__file__))
    src = json.load(open(os.path.join(os.path.dirname(here), "tests", "golden", "pyoracle_vectors.json")))
    doc = {"generator": "oracle/mporacle.py (mpmath, 50 digits, nested dual numbers)", "digits": 30, "cases": []}
    for c in src["cases"]:
        eoss = [Barton2009(b) for b in c["eos_blocks"]]
        Q = [mpf(float(x)) for x in c["Q"]]     # the FP64 state, taken as exact
        out = {"eos": c["eos"], "eos_blocks": c["eos_blocks"], "Q": c["Q"]}
        out["cons2prim"] = [s30(x) for p in range(2) for x in cons2prim(eoss[p], Q[15 * p:15 * p + 15])]
        out["flux"] = [s30(x) for p in range(2) for x in flux(eoss[p], Q[15 * p:15 * p + 15])]
        eg, acs = [], []
        for p in range(2):
            e, ac = get_eigvals_phase(eoss[p], Q[15 * p:15 * p + 15])
            eg += [s30(x) for x in e]; acs.append([[s30(x) for x in row] for row in ac])
        out["eigvals"] = eg
        out["acoustic"] = acs
        out["noncons_cols"] = [s30(x) for x in noncons_cols(eoss, Q)]
        doc["cases"].append(out)
    # one path-conservative HLL face per EoS set: between the first two states of the set
    faces = []
    for name in ("default", "hetero"):
        cs = [c for c in src["cases"] if c["eos"] == name]
        eoss = [Barton2009(b) for b in cs[0]["eos_blocks"]]
        Ql = [mpf(float(x)) for x in cs[0]["Q"]]; Qr = [mpf(float(x)) for x in cs[1]["Q"]]
        el = [x for p in range(2) for x in get_eigvals_phase(eoss[p], Ql[15 * p:15 * p + 15])[0]]
        er = [x for p in range(2) for x in get_eigvals_phase(eoss[p], Qr[15 * p:15 * p + 15])[0]]
        s_l, s_r, dm, dp = hll_pathcons(eoss, Ql, Qr, el, er)
        faces.append({"eos": name, "eos_blocks": cs[0]["eos_blocks"], "Ql": cs[0]["Q"], "Qr": cs[1]["Q"], "s_l": s30(s_l), "s_r": s30(s_r),
                      "dm": [s30(x) for x in dm], "dp": [s30(x) for x in dp]})
    doc["hll_faces"] = faces
    xs, ws = gauss_legendre6()
    doc["gauss_legendre6"] = {"x": [s30(x) for x in xs], "w": [s30(w) for w in ws]}
    # two complete time steps (main.jl:204-227) of a 10-cell Riemann grid: test case 6 states with a smooth
    # transition so that every face has a non-trivial path integral
    cs = [c for c in src["cases"] if c["eos"] == "default"]
    eoss = [Barton2009(b) for b in cs[0]["eos_blocks"]]
    Ql = [mpf(float(x)) for x in cs[0]["Q"]]; Qr = [mpf(float(x)) for x in cs[1]["Q"]]
    nx = 10
    cells = []
    for i in range(nx):
        w = mpf(i) / (nx - 1)
        w = float(w * w * (3 - 2 * w))                       # smoothstep, rounded to a double so that the FP64 codes get the same input
        cells.append([mpf(float((1 - w) * float(a) + w * float(b))) for a, b in zip(Ql, Qr)])
    start = [[float(x) for x in c] for c in cells]
    dts = []
    for _ in range(2):
        cells, dt = time_step(eoss, cells, mpf(float(0.6)), mpf(float(1.0 / nx)))
        dts.append(dt)
    doc["two_steps"] = {"eos_blocks": cs[0]["eos_blocks"], "nx": nx, "cfl": 0.6, "Q0": start, "dt": [s30(x) for x in dts],
                        "Q": [[s30(x) for x in c] for c in cells]}
    with open(path, "w") as f:
        json.dump(doc, f, indent=0)
    return doc


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    p = os.path.join(os.path.dirname(here), "tests", "golden", "mp_vectors.json")
    import time
    t0 = time.time()
    d = generate(p)
    print("wrote", p, len(d["cases"]), "cases,", len(d["hll_faces"]), "faces in", round(time.time() - t0, 1), "s")
