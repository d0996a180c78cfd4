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


@pytest.mark.parametrize("name", ["cas66", "cas75_odd", "one_beta_three_alpha", "gas343_54", "cas66_d2h", "gas2_sym4"])
def test_spin2_matrix_and_spin_adapted_guess(name):
    """f2_spin2_matrix (det.spin2 of the reference, determinant_helpers.cc:154-203): over a whole space its eigenvalues
    are S(S+1) with S = |Ms|, |Ms|+1, ..., it commutes with the Slater-rules Hamiltonian (which fixes every phase), and
    the guess made from it (ci_solver.spin_adapted_guess, forte2/sci/sci.py:405-455) is spin pure."""
    from forte2_b200.ci_solver import spin2_matrix, spin_adapted_guess, spin_complete
    space = SMALL_SPACES[name]
    s = f2.CIStrings(*space)
    E, H, V = synthetic.integrals(s.norb, "D")
    idx = np.arange(s.ndet, dtype=np.int64)
    S2, Hd = spin2_matrix(s, idx), slater_matrix(s, E, H, V, idx)
    assert np.allclose(S2, S2.T)
    assert np.abs(S2 @ Hd - Hd @ S2).max() < 1e-12 * max(1.0, np.abs(Hd).max())
    w = np.linalg.eigvalsh(S2)
    S = 0.5 * (-1.0 + np.sqrt(1.0 + 4.0 * np.maximum(w, 0.0)))
    ms = 0.5 * abs(space[0] - space[1])
    assert np.allclose(S - ms, np.round(S - ms), atol=1e-9) and S.min() > ms - 1e-9
    # guesses from a few determinants: spin-complete set, eigenvectors of S^2, wrong-spin vectors orthogonal to them
    mult = int(round(2 * ms)) + 1
    start = idx[:: max(1, s.ndet // 7)][:6]
    gidx, guess, wrong = spin_adapted_guess(s, E, H, V, start, mult, 4)
    assert set(start) <= set(gidx) and set(spin_complete(s, gidx)) == set(gidx)
    S2g = spin2_matrix(s, gidx)
    assert np.allclose(S2g @ guess, ms * (ms + 1) * guess, atol=1e-10)
    assert np.allclose(guess.T @ guess, np.eye(guess.shape[1]), atol=1e-10)
    assert np.abs(wrong.T @ guess).max(initial=0.0) < 1e-10
