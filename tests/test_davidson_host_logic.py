"""CPU tests of the Davidson-Liu iteration logic (forte2_b200/davidson.py) with the four streaming
kernels replaced by the host stand-in of tests/host_dl_kernels.py: same control flow as on the GPU,
checked against exact diagonalisation and, where /root/reference is present, against the
reference's own DavidsonLiuSolver (oracle/refdl.py) on the same matrix and guesses."""
import numpy as np
import pytest
import torch

from forte2_b200.davidson import DavidsonLiuParams, DavidsonLiuSolver
from host_dl_kernels import HostKernels


def _matrix(n, seed, spread=20.0, coupling=0.02):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(n, n))
    return coupling * (A + A.T) + np.diag(np.linspace(0.0, spread, n))


def _solve(A, nroot, params=None, guesses=None):
    n = A.shape[0]
    At = torch.from_numpy(A)
    s = DavidsonLiuSolver(size=n, nroot=nroot, davidson_liu_params=params, _kernels=HostKernels())
    s.add_h_diag(np.diag(A).copy())
    s.add_sigma_builder(lambda B, S: S.copy_(B @ At))
    if guesses is not None:
        s.add_guesses(guesses)
    return s, s.solve()


@pytest.mark.parametrize("nroot", [1, 2, 4])
def test_roots_match_eigh(nroot):
    A = _matrix(400, nroot)
    g = np.zeros((400, 2 * nroot))
    g[np.arange(2 * nroot), np.arange(2 * nroot)] = 1.0
    s, (lam, ev) = _solve(A, nroot, guesses=g)
    assert s.converged
    assert np.max(np.abs(lam - np.linalg.eigvalsh(A)[:nroot])) < 1e-11
    v = ev.numpy()
    assert np.allclose(v @ v.T, np.eye(nroot), atol=1e-10)


def test_random_guess_and_small_space():
    A = _matrix(7, 5, spread=3.0, coupling=0.3)
    s, (lam, _) = _solve(A, 2)  # no guesses: random start; basis saturates the 7-dimensional space
    assert np.max(np.abs(lam - np.linalg.eigvalsh(A)[:2])) < 1e-10


def test_collapse_keeps_the_ritz_values():
    A = _matrix(300, 9, coupling=0.2)
    p = DavidsonLiuParams(collapse_per_root=1, basis_per_root=2, maxiter=400)
    s, (lam, _) = _solve(A, 3, params=p)
    assert s.converged and np.max(np.abs(lam - np.linalg.eigvalsh(A)[:3])) < 1e-10


def test_energy_shift_targets_interior_roots():
    A = _matrix(200, 2, coupling=0.01)
    w = np.linalg.eigvalsh(A)
    n = A.shape[0]
    s = DavidsonLiuSolver(size=n, nroot=1, energy_shift=float(w[50]) + 1e-3, _kernels=HostKernels())
    s.add_h_diag(np.diag(A).copy())
    At = torch.from_numpy(A)
    s.add_sigma_builder(lambda B, S: S.copy_(B @ At))
    g = np.zeros((n, 1))
    g[50, 0] = 1.0
    s.add_guesses(g)
    lam, _ = s.solve()
    assert abs(lam[0] - w[50]) < 1e-9


def test_same_roots_as_reference_solver():
    from oracle import refdl
    if not refdl.available():
        pytest.skip("/root/reference not present")
    Solver, Params = refdl.load()
    A = _matrix(500, 21, coupling=0.05)
    g = np.zeros((500, 6))
    g[np.arange(6), np.arange(6)] = 1.0
    r = Solver(size=500, nroot=3, davidson_liu_params=Params(), log_level=5)
    r.add_h_diag(np.diag(A).copy())
    r.add_sigma_builder(lambda B, S: S.__setitem__(slice(None), A @ B))
    r.add_guesses(g)
    want, _ = r.solve()
    s, (lam, _) = _solve(A, 3, guesses=g)
    assert np.max(np.abs(lam - want)) < 1e-10
    # the first iterations (before any vector is discarded) follow the reference step by step
    assert s.history[0][4] == 9 and abs(s.history[0][1] - np.mean(np.linalg.eigvalsh(A[:6, :6])[:3])) < 1e-12


def test_chunked_ops_equal_the_unchunked_ones():
    """forte2_b200.davidson._ChunkedOps (the layer that cuts combine / residual into the kernels' 16 x 16 calls)
    against the same operations done in one piece: pure scaling of more than 16 rows, an in-place collapse whose
    outputs alias its inputs, accumulation into one vector from 40, residuals with more than 16 basis vectors
    and / or roots."""
    from host_dl_kernels import HostKernels16
    k, one = HostKernels16(), HostKernels()
    rng = np.random.default_rng(0)
    t = lambda *shape: torch.from_numpy(rng.normal(size=shape))  # noqa: E731
    Y = t(20, 7)
    Y0 = Y.clone()
    ss = k.combine(None, np.zeros((1, 20)), Y, beta=2.0, want_sumsq=True)
    assert torch.allclose(Y, 2 * Y0) and np.allclose(ss, (Y * Y).sum(1).numpy())
    A = t(40, 9)
    A0, c = A.clone(), rng.normal(size=(40, 20))
    k.combine(A, c, A[:20])  # outputs are the first 20 input rows
    assert torch.allclose(A[:20], torch.from_numpy(c).T @ A0)
    y = t(1, 9)
    y0, c1 = y.clone(), rng.normal(size=(40, 1))
    ss = k.combine(A0, c1, y, beta=0.5, want_sumsq=True)
    want = 0.5 * y0 + torch.from_numpy(c1).T @ A0
    assert torch.allclose(y, want) and np.allclose(ss, (want * want).sum(1).numpy())
    B, S, hd = t(40, 9), t(40, 9), t(9)
    al, lam = rng.normal(size=(40, 20)), rng.normal(size=20)
    for m in (40, 10):
        R, R1 = torch.zeros(20, 9, dtype=torch.float64), torch.zeros(20, 9, dtype=torch.float64)
        rn = k.residual(B[:m], S[:m], al[:m], lam, hd, R, 1e-8)
        rn1 = one.residual(B[:m], S[:m], al[:m], lam, hd, R1, 1e-8)
        assert torch.allclose(R, R1) and np.allclose(rn, rn1)


def test_block_orthonormalisation_is_the_path_taken_and_handles_dependent_candidates():
    """at most 4 roots and 16 basis vectors: the corrections are orthonormalised as a block (two passes over the
    basis per iteration instead of 3 per candidate); candidates that are nearly dependent on the basis or on each
    other go through the re-projection / refinement cycles and still give an orthonormal basis"""
    calls = {"project": 0, "residual_fused": 0}

    class Counting(HostKernels):
        def project(self, *a):
            calls["project"] += 1
            return HostKernels.project(self, *a)

        def residual_fused(self, *a):
            calls["residual_fused"] += 1
            return HostKernels.residual_fused(self, *a)

    A = _matrix(500, 21, coupling=0.05)
    At = torch.from_numpy(A)
    s = DavidsonLiuSolver(size=500, nroot=3, _kernels=Counting())
    s.add_h_diag(np.diag(A).copy())
    s.add_sigma_builder(lambda B, S: S.copy_(B @ At))
    lam, ev = s.solve()
    assert s.converged and np.max(np.abs(lam - np.linalg.eigvalsh(A)[:3])) < 1e-10
    assert calls["residual_fused"] == len(s.iter_times) and calls["project"] >= 2 * (len(s.iter_times) - 1)
    assert len(s.iter_bytes) == len(s.iter_times)
    # direct: a block with one candidate inside the span of the basis, one nearly parallel to another
    rng = np.random.default_rng(4)
    s2 = DavidsonLiuSolver(size=300, nroot=3, _kernels=HostKernels())
    Q, _ = np.linalg.qr(rng.normal(size=(300, 5)))
    s2.b[:5] = torch.from_numpy(Q.T.copy())
    s2.basis_size = 5
    v = rng.normal(size=300)
    cand = torch.from_numpy(np.stack([Q[:, 0] * 2.0 + 1e-12 * v, v, v * (1 + 1e-9) + 1e-7 * rng.normal(size=300)]))
    added = s2._add_block_orthonormalized(cand, None)
    assert added == 2 and s2._accepted == [False, True, True]
    Bn = s2.b[:7].numpy()
    assert np.max(np.abs(Bn @ Bn.T - np.eye(7))) < 1e-12
