"""GPU checks of code written after this round's GPU budget was spent: they have not run on hardware yet.  Each
one exercises host-side glue around GPU-validated kernels (strided numpy views in the spin adapter, the host Slater-rule
helpers of the builder, forte2's sigma closure replayed verbatim, ci_algorithm="exact", a Davidson subspace beyond the
kernels' 16 vectors per call) whose logic is covered by the CPU suite against the reference.  They sort last and are
marked xfail(strict=False): the round-end run records pass (XPASS) or fail without gating the validated suite.
Once seen passing they belong in test_spin_adapter.py / test_davidson_gpu.py without the marker."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
import forte2_b200 as f2
from forte2_b200 import synthetic

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(strict=False, reason="first run on hardware (written without GPU access)")]
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SPACES = SMALL_SPACES
ROOT_TOL = 1e-10
CASES = sorted((f[4:-4].rsplit("_S", 1)[0], int(f[4:-4].rsplit("_S", 1)[1])) for f in os.listdir(GOLD)
               if f.startswith("csf_"))


@pytest.fixture
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _adapter(name, twoS):
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    a = f2.CISpinAdapter(twoS, space[0] - space[1], strings.norb)
    a.prepare_couplings(strings)
    return strings, a


@pytest.mark.parametrize("name,twoS", CASES)
def test_adapter_accepts_strided_numpy_views(torch_cuda, name, twoS):
    g = np.load(os.path.join(GOLD, f"csf_{name}_S{twoS}.npz"))
    strings, a = _adapter(name, twoS)
    to_det = np.full(strings.ndet, np.nan)
    a.csf_C_to_det_C(g["c_csf"], to_det)
    to_csf = np.full(a.ncsf, np.nan)
    a.det_C_to_csf_C(g["c_det"], to_csf)
    # columns of a C-ordered (size, m) basis, as forte2's Davidson closure passes them (ci.py:231-233)
    basis = np.full((a.ncsf, 3), np.nan)
    basis[:, 1] = g["c_csf"]
    to_det2 = np.full(strings.ndet, np.nan)
    a.csf_C_to_det_C(basis[:, 1], to_det2)
    assert np.max(np.abs(to_det2 - to_det), initial=0.0) <= 1e-14
    out = np.full((a.ncsf, 3), np.nan)
    a.det_C_to_csf_C(g["c_det"], out[:, 2])
    assert np.max(np.abs(out[:, 2] - to_csf), initial=0.0) <= 1e-14 and np.all(np.isnan(out[:, :2]))


@pytest.mark.parametrize("name,twoS", CASES)
def test_builder_csf_helpers_and_forte2_closure(torch_cuda, name, twoS):
    strings, a = _adapter(name, twoS)
    E, H, V = synthetic.integrals(strings.norb, "S")
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    # the exact form: host Slater rules (checked against the reference on the CPU), delivered as a device tensor
    from forte2_b200.ci_sigma_builder import exact_hdiag_csf
    hx = b.form_Hdiag_csf(None, a, True)
    assert hx.is_cuda and hx.shape == (a.ncsf,)
    assert np.array_equal(hx.cpu().numpy(), exact_hdiag_csf(strings, E, H, V, a))
    # host helpers of the guess / exact diagonalisation through the builder (checked against the reference on the CPU)
    from forte2_b200.ci_sigma_builder import csf_hamiltonian, csf_matrix_element
    assert b.slater_rules_csf(None, a, 0, a.ncsf - 1) == csf_matrix_element(strings, E, H, V, a, 0, a.ncsf - 1)
    assert b.energy_csf(None, a, 0) == csf_matrix_element(strings, E, H, V, a, 0, 0)
    if a.ncsf <= 400:
        Hc = b.form_H_csf(None, a)
        assert np.array_equal(Hc, csf_hamiltonian(strings, E, H, V, a))
        assert np.max(np.abs(np.diag(Hc) - hx.cpu().numpy())) <= 1e-12 * max(1.0, np.max(np.abs(Hc)))
        # forte2's sigma closure, verbatim (forte2/ci/ci.py:221-233): numpy vectors, columns of C-ordered blocks
        rng = np.random.default_rng(5)
        Bblock = rng.normal(size=(a.ncsf, 2))
        Sblock = np.full_like(Bblock, np.nan)
        b_det, sigma_det = np.zeros(strings.ndet), np.zeros(strings.ndet)
        for i in range(Bblock.shape[1]):
            a.csf_C_to_det_C(Bblock[:, i], b_det)
            b.Hamiltonian(b_det, sigma_det)
            a.det_C_to_csf_C(sigma_det, Sblock[:, i])
        assert np.max(np.abs(Sblock - Hc @ Bblock)) <= 1e-12 * max(1.0, np.max(np.abs(Hc)))
        b.set_memory(0)
        with pytest.raises(RuntimeError, match="Not enough memory"):
            b.form_H_csf(None, a)


def test_solver_with_24_subspace_vectors(torch_cuda):
    """6 roots x basis_per_root 4: combine / residual go through the 16 x 16 chunked calls (davidson._ChunkedOps)"""
    torch = torch_cuda
    from forte2_b200.davidson import DavidsonLiuSolver
    n, nroot = 600, 6
    rng = np.random.default_rng(3)
    A = rng.normal(size=(n, n))
    A = 0.02 * (A + A.T) + np.diag(np.linspace(0.0, 20.0, n))
    Ad = torch.from_numpy(A).cuda()
    s = DavidsonLiuSolver(size=n, nroot=nroot)
    s.add_h_diag(np.diag(A).copy())
    s.add_sigma_builder(lambda Bb, Sb: Sb.copy_(Bb @ Ad))
    g = np.zeros((n, 2 * nroot))
    g[np.arange(2 * nroot), np.arange(2 * nroot)] = 1.0
    s.add_guesses(g)
    lam, ev = s.solve()
    assert s.converged
    want = np.linalg.eigvalsh(A)[:nroot]
    assert np.max(np.abs(lam - want)) < 1e-11
    evh = ev.cpu().numpy()
    assert np.allclose(evh @ evh.T, np.eye(nroot), atol=1e-10)
    assert np.max(np.abs(evh @ A - lam[:, None] * evh)) < 1e-5


@pytest.mark.parametrize("name", ["cas66", "cas66_d2h"])
def test_exact_diagonalization_equals_the_iterative_solve(torch_cuda, name):
    """ci_algorithm="exact" (forte2/ci/ci.py:305-322; tests/ci/test_ci_exact_diagonalization.py pins HZ == exact):
    dense Hamiltonian from host Slater rules, in the determinant basis and in the CSF basis, against the committed
    exact eigenvalues, and its eigenvector against one sigma build."""
    import forte2_b200 as f2
    from forte2_b200.ci_solver import CISolver
    torch = torch_cuda
    space = SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "D")
    strings = f2.CIStrings(*space)
    g = np.load(os.path.join(GOLD, f"roots_{name}_D.npz"))
    solver = CISolver(strings, E, H, V, nroot=3, ci_algorithm="exact", log_level=0).run()
    assert solver.converged and solver.evecs.is_cuda and solver.evecs.shape == (3, strings.ndet)
    assert np.max(np.abs(solver.E - g["exact"][:3])) <= ROOT_TOL
    sg = torch.empty_like(solver.evecs[0])
    solver.ci_sigma_builder.Hamiltonian(solver.evecs[0].contiguous(), sg)
    assert float((sg - solver.E[0] * solver.evecs[0]).norm()) < 1e-9
    # CSF basis (singlets): the lowest singlet of the dense CSF Hamiltonian is a root of the determinant problem
    gc = np.load(os.path.join(GOLD, f"rootscsf_{name}_S0.npz"))
    adapter = f2.CISpinAdapter(0, space[0] - space[1], n)
    adapter.prepare_couplings(strings)
    csf = CISolver(strings, E, H, V, nroot=2, ci_algorithm="exact", log_level=0, spin_adapter=adapter).run()
    assert csf.evecs.shape == (2, adapter.ncsf)
    assert np.max(np.abs(csf.E - gc["exact"][:2])) <= ROOT_TOL
    shifted = CISolver(strings, E, H, V, nroot=1, ci_algorithm="exact", log_level=0, spin_adapter=adapter,
                       energy_shift=float(gc["exact"][1]) + 1e-6).run()
    assert abs(shifted.E[0] - gc["exact"][1]) <= ROOT_TOL
