"""Host Slater rules of the C ABI (f2_slater_matrix, no GPU needed) against the reference's dense
Hamiltonian (oracle/_ref when built) and against the exact spectra stored in the golden fixtures."""
import os

import numpy as np
import pytest

import forte2_b200 as f2
from cases import SMALL_SPACES
from forte2_b200 import synthetic
from forte2_b200.ci_solver import slater_matrix

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name,variant", [("cas66", "S"), ("cas66_d2h", "S"), ("cas21_4", "four_fold"),
                                          ("sym2_11", "S"), ("gas2_sym4", "four_fold"), ("one_each", "S")])
def test_spectrum_matches_golden(name, variant):
    g = np.load(os.path.join(GOLD, f"sigma_{name}_{variant}.npz"))
    if "evals" not in g.files:
        pytest.skip("space too large for a stored spectrum")
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.four_fold_integrals(n) if variant == "four_fold" else synthetic.integrals(n, variant)
    s = f2.CIStrings(*space)
    Hd = slater_matrix(s, E, H, V, np.arange(s.ndet))
    assert np.allclose(Hd, Hd.T)
    assert np.allclose(np.diag(Hd), g["hdiag"], rtol=0, atol=1e-12)
    assert np.allclose(Hd @ g["c"], g["sigma_hz"], rtol=0, atol=1e-12)
    assert np.allclose(np.linalg.eigvalsh(Hd)[: len(g["evals"])], g["evals"], rtol=0, atol=1e-11)


@pytest.mark.parametrize("name", ["gas343_54", "cas75_odd", "beta_only", "gas3_53"])
def test_matches_live_reference(ref, name):
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.four_fold_integrals(n)
    rb = ref.RefCISigmaBuilder(ref.RefCIStrings(*space), E, H, V)
    want = rb.dense_hamiltonian()
    s = f2.CIStrings(*space)
    rng = np.random.default_rng(1)
    idx = rng.permutation(s.ndet)[: min(s.ndet, 200)]
    got = slater_matrix(s, E, H, V, idx)
    assert np.max(np.abs(got - want[np.ix_(idx, idx)])) <= 1e-13


def test_address_out_of_range():
    s = f2.CIStrings(2, 2, 0, [[0] * 4], [], [])
    with pytest.raises(IndexError):
        slater_matrix(s, 0.0, np.zeros((4, 4)), np.zeros((4,) * 4), [s.ndet])
