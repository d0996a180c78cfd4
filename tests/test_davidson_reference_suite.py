"""The reference's own Davidson-Liu tests (tests/helpers/test_davidsonliu.py: test_dl_1..4, test_dl_no_guess,
test_dl_restart_1/2, test_dl_restart_from_unconverged) restated for forte2_b200.davidson.DavidsonLiuSolver with
the streaming kernels replaced by the host stand-in of tests/host_dl_kernels.py (behind the product's own
16 x 16 chunking layer, so subspaces of up to 50 vectors run the decomposition the GPU path uses): same matrices, same parameters,
same assertions (the reference's `approx` is abs 5e-8, forte2/helpers/comparisons.py)."""
import numpy as np
import pytest
import torch

from forte2_b200.davidson import DavidsonLiuParams, DavidsonLiuSolver
from host_dl_kernels import HostKernels16

ABS = 5e-8
M4 = np.array([[-1, 1, 1, 1], [1, 0, 1, 1], [1, 1, 0, 1], [1, 1, 1, 0]], dtype=np.float64)


def _solver(matrix, nroot, params):
    s = DavidsonLiuSolver(size=matrix.shape[0], nroot=nroot, davidson_liu_params=params, log_level=5,
                          _kernels=HostKernels16())
    s.add_h_diag(np.diag(matrix).copy())
    At = torch.from_numpy(np.ascontiguousarray(matrix))
    s.add_sigma_builder(lambda B, S: S.copy_(B @ At))
    return s


def _p15(**kw):
    return DavidsonLiuParams(collapse_per_root=1, basis_per_root=5, **kw)


def test_dl_1_standard_basis_guesses():  # test_davidsonliu.py:91-123
    s = _solver(M4, 1, _p15())
    s.add_guesses(np.eye(4))
    evals, _ = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(M4)[0]) <= ABS


def test_dl_2_one_unnormalised_guess():  # :126-157
    s = _solver(M4, 1, _p15())
    s.add_guesses(np.array([[3.0], [0.0], [0.0], [0.0]]))
    evals, evecs = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(M4)[0]) <= ABS
    v = np.asarray(evecs).reshape(-1)
    assert abs(np.linalg.norm(v) - 1.0) <= 1e-10


def _banded(size):  # solve_dl, :160-199
    m = np.zeros((size, size))
    for i in range(size):
        m[i, i] = -1.0 + i * 0.1
        for j in range(i):
            m[i, j] = m[j, i] = 0.05 / (1.0 + abs(i - j))
    return m


def _solve_dl(size, nroot):
    m = _banded(size)
    s = _solver(m, nroot, _p15())
    s.add_guesses(np.eye(size)[:, :nroot])
    evals, _ = s.solve()
    assert np.allclose(np.sort(evals), np.linalg.eigvalsh(m)[:nroot], atol=1e-6), (size, nroot, evals)


@pytest.mark.parametrize("size", [10, 100, 1000])
def test_dl_3_sizes_and_roots(size):  # :202-207 (nroot 1..10)
    for nroot in (1, 2, 3, 5, 10):
        _solve_dl(size, nroot)


def test_dl_4_every_root_count_of_small_matrices():  # :210-214 (the reference runs every size below 40)
    for size in (1, 2, 3, 4, 5, 7, 12, 23, 39):
        for nroot in sorted({1, 2, size // 2, size - 1, size} - {0}):
            if nroot <= size:
                _solve_dl(size, nroot)


def test_dl_no_guess():  # :217-239
    s = _solver(M4, 1, _p15())
    evals, _ = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(M4)[0]) <= ABS


def test_dl_restart_same_matrix():  # :269-293
    s = _solver(M4, 1, _p15())
    s.add_guesses(np.eye(4)[:, :1])
    s.solve()
    evals, _ = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(M4)[0]) <= ABS


def test_dl_restart_changed_matrix():  # :296-330
    s = _solver(M4, 1, _p15())
    s.add_guesses(np.eye(4)[:, :1])
    evals, _ = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(M4)[0]) <= ABS
    m2 = M4.copy()
    m2[0, 0] = -2.0
    A2 = torch.from_numpy(m2)
    s.add_h_diag(np.diag(m2).copy())
    s.add_sigma_builder(lambda B, S: S.copy_(B @ A2))
    evals, _ = s.solve()
    assert abs(evals[0] - np.linalg.eigvalsh(m2)[0]) <= ABS


def test_dl_restart_from_unconverged():  # :333-361
    size, nroot = 500, 10
    rng = np.random.default_rng(42)
    m = rng.random((size, size))
    m = 0.05 * (m + m.T) + np.diag(rng.uniform(2, 4, size))
    s = _solver(m, nroot, DavidsonLiuParams(collapse_per_root=2, basis_per_root=5, maxiter=10))
    s.add_guesses(np.eye(size)[:, :nroot])
    evals, _ = s.solve()
    rounds = 1
    while not s.converged:
        evals, _ = s.solve()
        rounds += 1
        assert rounds < 50
    assert np.max(np.abs(np.asarray(evals)[:nroot] - np.linalg.eigvalsh(m)[:nroot])) <= ABS
