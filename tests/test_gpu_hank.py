"""GPU: the Hank2016 batches (hs_hank2016_energy / pressure / stress, EquationsOfState.jl:301-356)
through the C ABI against the literal dual-number oracle, <= 1e-12 relative per evaluation."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cases(oracle, rng, n):
    A = np.eye(3)[None] + 0.2 * rng.uniform(-1, 1, (n, 3, 3))
    a9 = np.ascontiguousarray(A.transpose(0, 2, 1).reshape(n, 9))            # column-major entries
    G = np.einsum("nki,nkj->nij", A, A)                                       # A^T A
    g9 = np.ascontiguousarray(G.transpose(0, 2, 1).reshape(n, 9))
    den = 2.7 * np.abs(np.linalg.det(A)) * rng.uniform(0.9, 1.1, n)
    pres = rng.uniform(-1e9, 5e10, n)
    inv3 = np.stack([oracle.invariants(g) for g in g9])
    return a9, g9, inv3, den, pres


@pytest.mark.parametrize("kw", [{}, dict(rho0=8.9, mu=48e9, gamma=4.2, pres_inf=34e9, a=-0.3)])
def test_hank2016_batches_vs_oracle(gpu, oracle, kw):
    E = gpu.equations_of_state
    eos, eo = gpu.Hank2016(**kw), oracle.hank2016(**kw)
    rng = np.random.default_rng(5)
    n = 1000                                                                   # 7 full blocks + a ragged one
    a9, g9, inv3, den, pres = _cases(oracle, rng, n)
    e = E.energy(eos, den, pres, g9)
    e_ref = np.array([oracle.hank_energy(eo, den[i], pres[i], g9[i])[0] for i in range(n)])
    assert np.abs(e - e_ref).max() <= 1e-13 * np.abs(e_ref).max()
    p = E.pressure(eos, den, e_ref, inv3)
    p_ref = np.array([oracle.hank_pressure(eo, den[i], e_ref[i], inv3[i])[0] for i in range(n)])
    scale = eo[2] * eo[3]
    assert np.abs(p - p_ref).max() <= 1e-12 * scale
    assert np.abs(p - pres).max() <= 1e-11 * scale                             # pressure inverts energy
    s = E.stress(eos, den, pres, a9)
    s_ref = np.stack([oracle.hank_stress(eo, den[i], pres[i], a9[i])[0] for i in range(n)])
    assert np.abs(s - s_ref).max() <= 1e-12 * np.abs(s_ref).max()
    # scalar methods, (3, 3) matrices as the reference passes them
    A0 = a9[0].reshape(3, 3).T
    assert abs(E.energy(eos, den[0], pres[0], A0.T @ A0) - e_ref[0]) <= 1e-13 * abs(e_ref[0])
    assert np.abs(E.stress(eos, den[0], pres[0], A0) - s_ref[0]).max() <= 1e-12 * np.abs(s_ref).max()
    assert abs(E.pressure(eos, den[0], e_ref[0], inv3[0]) - p_ref[0]) <= 1e-12 * scale


def test_hank2016_edge_cases(gpu, oracle):
    E = gpu.equations_of_state
    eos = E.eos_hank2016
    # undeformed state: no elastic energy, no stress (single item = a ragged block of one)
    e = E.energy(eos, 2.7, 1e9, np.eye(3))
    assert abs(e - (1e9 + 3.4 * 21.5e9) / (2.7 * 2.4)) <= 1e-15 * e
    assert np.abs(E.stress(eos, 2.7, 0.0, np.eye(3))).max() <= 1e-14 * 26e9
    # det G <= 0: Julia's fractional power throws DomainError
    with pytest.raises(gpu.DomainError):
        E.energy(eos, 2.7, 1e9, np.diag([1.0, 1.0, -1.0]))
    with pytest.raises(gpu.DomainError):
        E.pressure(eos, 2.7, 1e9, np.array([3.0, 3.0, -1.0]))
    # argument errors
    with pytest.raises(ValueError):
        E.energy(eos, np.ones(3), np.ones(2), np.tile(np.eye(3).reshape(1, 9), (3, 1)))
    with pytest.raises(TypeError):
        E.energy(gpu.Barton2009(), 1.0, 1.0, np.eye(3))
    rc = gpu.lib().hs_hank2016_energy(None, None, None, None, None, 1, 0)
    assert rc == 1                                                             # HS_ERR_ARG


def test_hank2016_golden_vectors(gpu):
    """the committed torch-autograd vectors (oracle/pyoracle.py::generate_hank) through the C ABI"""
    import json, os
    E = gpu.equations_of_state
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyoracle_hank_vectors.json")))
    for c in d["cases"]:
        eos = gpu.Hank2016(*c["eos_block"])
        assert abs(E.energy(eos, c["den"], c["pres"], np.array(c["G"])) - c["energy"]) <= 1e-13 * abs(c["energy"])
        assert abs(E.pressure(eos, c["den"], c["energy"], np.array(c["invariants"])) - c["pressure"]) <= 1e-12 * eos.gamma * eos.pres_inf
        s = E.stress(eos, c["den"], c["pres"], np.array(c["distortion"]))
        assert np.abs(s - np.array(c["stress"])).max() <= 1e-12 * np.abs(c["stress"]).max()
