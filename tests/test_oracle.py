"""The oracle checks itself: the C/Python restatement (oracle/sigma_port.c, oracle/strings_port.py)
against the committed golden vectors generated from the reference, and against the live reference
library (oracle/_ref) where it has been built."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
from forte2_b200 import synthetic
from oracle import checker, strings_port

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SIGMA_FIXTURES = sorted(f[len("sigma_"):-4] for f in os.listdir(GOLD) if f.startswith("sigma_"))


def _ints(n, variant):
    return synthetic.four_fold_integrals(n) if variant == "four_fold" else synthetic.integrals(n, variant)


@pytest.mark.parametrize("fixture", SIGMA_FIXTURES)
def test_port_matches_golden_sigma(fixture):
    name, variant = fixture.rsplit("_", 1) if not fixture.endswith("four_fold") else (fixture[:-10], "four_fold")
    space = SMALL_SPACES[name]
    g = np.load(os.path.join(GOLD, f"sigma_{fixture}.npz"))
    dets = strings_port.determinants(*space)
    assert np.array_equal(dets, g["dets"])  # determinant order restated exactly
    n = sum(len(s) for s in space[3])
    E, H, V = _ints(n, variant)
    got = checker.port_sigma(dets, E, H, V, g["c"])
    assert np.linalg.norm(got - g["sigma_hz"]) <= 1e-13 * max(1.0, np.linalg.norm(g["sigma_hz"]))
    assert np.allclose(checker.port_energies(dets, E, H, V), g["hdiag"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["cas66", "gas343_55", "cas66_d2h", "gas2_sym4", "cas75_odd"])
def test_port_matches_live_reference(ref, name):
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = _ints(n, "D")
    rs = ref.RefCIStrings(*space)
    dets = strings_port.determinants(*space)
    assert np.array_equal(dets, rs.determinants())
    c = synthetic.trial_vector(rs.ndet, seed=7)
    want, kind = checker.sigma(space, E, H, V, c)
    assert kind == "reference"
    got = checker.port_sigma(dets, E, H, V, c)
    assert np.linalg.norm(got - want) <= 1e-13 * max(1.0, np.linalg.norm(want))


def test_reference_hz_kh_dense_agree(ref):
    """the reference's own cross-check (tests/ci/test_gasci_rhf.py:222-276, test_ci_exact_diagonalization.py):
    HZ == KH == dense Slater rules, here on synthetic 8-fold symmetric integrals"""
    space = SMALL_SPACES["cas66"]
    E, H, V = _ints(6, "S")
    rs = ref.RefCIStrings(*space)
    b = ref.RefCISigmaBuilder(rs, E, H, V)
    c = synthetic.trial_vector(rs.ndet)
    out = {}
    for alg in ("hz", "kh"):
        b.set_algorithm(alg)
        out[alg] = np.zeros_like(c)
        b.Hamiltonian(c, out[alg])
    dense = b.dense_hamiltonian() @ c
    assert np.linalg.norm(out["hz"] - out["kh"]) <= 1e-13 * np.linalg.norm(dense)
    assert np.linalg.norm(out["hz"] - dense) <= 1e-13 * np.linalg.norm(dense)
