import numpy as np


def relerr(a, b, per_var=True):
    """max over cells of |a-b| scaled by the max-norm of each variable of b (parity metric:
    'relative tolerance on conserved variables')."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if per_var and a.ndim >= 2:
        ax = tuple(range(a.ndim - 1))
        scale = np.maximum(np.abs(b).max(axis=ax), 1e-300)
        # variables that are identically 0 up to roundoff (e.g. a momentum component of a state at rest) carry
        # only noise: scale them by the size of the state instead of by their own ~1e-16 magnitude
        scale = np.maximum(scale, 1e-3 * np.abs(b).max())
        return float((np.abs(a - b).max(axis=ax) / scale).max())
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def random_mph_prims(rng, n, eos_rho_nominal=8.9, spread=0.05, same_phases=True):
    """Random admissible primitive two-phase states (BASELINE.json config 4 generator)."""
    P = np.zeros((n, 30))
    for i in range(n):
        a1 = rng.uniform(0.1, 0.9)
        for p, a in enumerate((a1, 1 - a1)):
            if p == 0 or not same_phases:
                u = rng.uniform(-1, 1, 3); S = rng.uniform(0, 1e-3)
                F = np.eye(3) + spread * rng.uniform(-1, 1, (3, 3))
            rho = eos_rho_nominal / np.linalg.det(F)
            P[i, 15 * p:15 * p + 15] = [a, rho, *u, S, *F.flatten(order="F")]
    return P


def random_sp_prims(rng, n, spread=0.05):
    P = np.zeros((n, 13))
    for i in range(n):
        u = rng.uniform(-1, 1, 3); S = rng.uniform(0, 1e-3)
        F = np.eye(3) + spread * rng.uniform(-1, 1, (3, 3))
        P[i] = [*u, *F.flatten(order="C"), S]
    return P
