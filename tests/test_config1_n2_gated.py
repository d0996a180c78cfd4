"""BASELINE.json configs[0]: N2 / cc-pVDZ RHF -> CASCI(6e,6o), the reference's own CPU test path
(tests/ci/test_ci_rhf.py:57-78 `test_ci_n2_with_symmetry`, :107-134 `test_sa_ci_n2`).

The energies need the molecular integrals, which need libint2: not available in this image (SURVEY.md section
8c), so the integrals cannot be produced here.  These tests run when a fixture exported from a real forte2
install is present ("Exporting the config-1 integrals"):

    baseline/fixtures/n2_cas66.npz       E (scalar), H (6,6), V (6,6,6,6) physicist <pq|rs>: ci.ints.E/H/V of
                                         test_sa_ci_n2 (no symmetry)
    baseline/fixtures/n2_cas66_d2h.npz   the same of test_ci_n2_with_symmetry plus orbital_symmetry (6 irrep labels)

Without them they are skipped; config 1 is then covered on synthetic integrals (tests/test_sigma_gpu.py,
tests/test_davidson_gpu.py: CAS(6,6) with and without 8 irreps against the reference's builder and solver)."""
import os

import numpy as np
import pytest

import forte2_b200 as f2

pytestmark = pytest.mark.gpu
FIX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "fixtures")
ABS = 5e-8  # forte2/helpers/comparisons.py
E_SINGLET, E_TRIPLET1, E_TRIPLET2 = -109.004622061660, -108.779926502402, -108.733907910380


def _load(name):
    path = os.path.join(FIX, name)
    if not os.path.exists(path):
        pytest.skip(f"{path} not present (needs integrals from a forte2 install with libint2)")
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(path)


def _solve(orbital_symmetry, symmetry, twoS, nroot, g):
    from forte2_b200.ci_solver import CISolver
    strings = f2.CIStrings(3, 3, symmetry, [list(orbital_symmetry)], [], [])
    adapter = f2.CISpinAdapter(twoS, 0, 6)
    adapter.prepare_couplings(strings)
    solver = CISolver(strings, float(g["E"]), g["H"], g["V"], nroot=nroot, log_level=0, spin_adapter=adapter)
    return solver.run().E


def test_sa_ci_n2():
    g = _load("n2_cas66.npz")
    assert abs(_solve([0] * 6, 0, 0, 1, g)[0] - E_SINGLET) <= ABS
    triplets = _solve([0] * 6, 0, 2, 2, g)
    assert abs(triplets[0] - E_TRIPLET1) <= ABS and abs(triplets[1] - E_TRIPLET2) <= ABS


def test_ci_n2_with_symmetry():
    g = _load("n2_cas66_d2h.npz")
    assert abs(_solve(g["orbital_symmetry"].tolist(), 0, 0, 1, g)[0] - E_SINGLET) <= ABS
