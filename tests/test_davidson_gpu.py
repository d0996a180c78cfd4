"""GPU parity of the device Davidson-Liu path: the f2_dl_* streaming kernels against plain torch, the
solver against exact diagonalisation, and CISolver roots against the golden roots produced by the
reference's own builder + DavidsonLiuSolver (tests/golden/make_golden_roots.py).
Tolerance from BASELINE.json: converged roots <= 1e-10 Eh."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
from forte2_b200 import synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SPACES = dict(SMALL_SPACES, cas1010=(5, 5, 0, [[0] * 10], [], []))
ROOT_TOL = 1e-10


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.parametrize("n", [1, 7, 1000, 65537, 1 << 20])
def test_streaming_kernels_match_torch(torch_cuda, n):
    torch = torch_cuda
    from forte2_b200.davidson import _Kernels
    k = _Kernels(torch, torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(n)
    for p, q in [(1, 1), (3, 2), (5, 3), (12, 3), (12, 6), (16, 16)]:
        A = torch.randn((p, n), dtype=torch.float64, device="cuda", generator=g)
        X = torch.randn((q, n), dtype=torch.float64, device="cuda", generator=g)
        want = (A @ X.T).cpu().numpy()
        got = k.dots(A, X)
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12 * np.sqrt(n))
        # combine, out of place with beta, and the sum of squares
        coef = np.random.default_rng(p * 31 + q).normal(size=(p, q))
        Y = torch.randn((q, n), dtype=torch.float64, device="cuda", generator=g)
        wantY = 0.5 * Y + torch.from_numpy(coef).cuda().T @ A
        ss = k.combine(A, coef, Y, beta=0.5, want_sumsq=True)
        assert torch.allclose(Y, wantY, rtol=1e-13, atol=1e-13)
        assert np.allclose(ss, (wantY * wantY).sum(dim=1).cpu().numpy(), rtol=1e-12)
        # in place: the first q rows of A are overwritten by combinations of all p rows
        if q <= p:
            A2 = A.clone()
            k.combine(A2, coef, A2[:q])
            assert torch.allclose(A2[:q], torch.from_numpy(coef).cuda().T @ A, rtol=1e-13, atol=1e-13)
            assert torch.equal(A2[q:], A[q:])
    # residual + preconditioner
    m, kr = 7, 3
    B = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    S = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    h = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    lam = np.array([0.1, -0.3, 0.7])
    h[:: 5] = lam[0] + 1e-9  # small denominators are masked to zero
    alpha = np.random.default_rng(5).normal(size=(m, kr))
    R = torch.empty((kr, n), dtype=torch.float64, device="cuda")
    rn2 = k.residual(B, S, alpha, lam, h, R, 1e-8)
    a = torch.from_numpy(alpha).cuda()
    Rw = a.T @ S - torch.from_numpy(lam).cuda()[:, None] * (a.T @ B)
    assert np.allclose(rn2, (Rw * Rw).sum(dim=1).cpu().numpy(), rtol=1e-12)
    den = torch.from_numpy(lam).cuda()[:, None] - h[None, :]
    Rp = torch.where(den.abs() > 1e-8, Rw / den, torch.zeros_like(Rw))
    assert torch.allclose(R, Rp, rtol=1e-12, atol=1e-12)
    v = B[0].clone()
    k.scale(v, -2.5)
    assert torch.allclose(v, -2.5 * B[0])


@pytest.mark.parametrize("n", [1, 1000, 65537, (1 << 20) + 3])
def test_fused_passes_match_torch(torch_cuda, n):
    """f2_dl_residual_fused / f2_dl_project (block orthonormalisation of the candidates, csrc/davidson.cu) against
    plain fp64 torch: vectors written, and the overlaps returned from the same pass"""
    torch = torch_cuda
    from forte2_b200.davidson import _Kernels
    k = _Kernels(torch, torch.device("cuda", 0))
    g = torch.Generator(device="cuda").manual_seed(n + 1)
    tol = dict(rtol=1e-11, atol=1e-11 * np.sqrt(n))
    for m, kr in [(1, 1), (3, 2), (6, 3), (12, 3), (9, 4), (16, 4)]:
        B = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
        S = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
        h = torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
        lam = np.linspace(-0.3, 0.7, kr)
        h[:: 5] = lam[0] + 1e-9
        alpha = np.random.default_rng(m * 7 + kr).normal(size=(m, kr))
        R = torch.full((kr, n), float("nan"), dtype=torch.float64, device="cuda")
        rn2, c, gg = k.residual_fused(B, S, alpha, lam, h, R, 1e-8)
        a = torch.from_numpy(alpha).cuda()
        Rw = a.T @ S - torch.from_numpy(lam).cuda()[:, None] * (a.T @ B)
        den = torch.from_numpy(lam).cuda()[:, None] - h[None, :]
        Rp = torch.where(den.abs() > 1e-8, Rw / den, torch.zeros_like(Rw))
        assert np.allclose(rn2, (Rw * Rw).sum(dim=1).cpu().numpy(), rtol=1e-12)
        assert torch.allclose(R, Rp, rtol=1e-12, atol=1e-12)
        scale = max(1.0, float(Rp.abs().max()))
        assert np.allclose(c, (B @ Rp.T).cpu().numpy(), rtol=1e-11, atol=1e-11 * np.sqrt(n) * scale)
        assert np.allclose(gg, (Rp @ Rp.T).cpu().numpy(), rtol=1e-11, atol=1e-11 * np.sqrt(n) * scale ** 2)
        # projection + in-block transform, fewer outputs than inputs, out of place and in place
        for kout in sorted({1, kr}):
            Y = torch.randn((kr, n), dtype=torch.float64, device="cuda", generator=g)
            cc = np.random.default_rng(kout).normal(size=(m, kr))
            T = np.random.default_rng(kout + 9).normal(size=(kr, kout))
            want = torch.from_numpy(T).cuda().T @ (Y - torch.from_numpy(cc).cuda().T @ B)
            Z = torch.full((kout, n), float("nan"), dtype=torch.float64, device="cuda")
            d, g2 = k.project(B, Y, cc, T, Z)
            assert torch.allclose(Z, want, rtol=1e-12, atol=1e-12)
            assert np.allclose(d, (B @ want.T).cpu().numpy(), **tol) and np.allclose(g2, (want @ want.T).cpu().numpy(), **tol)
            Y2 = Y.clone()
            d, g2 = k.project(B, Y2, cc, T, Y2[:kout])
            assert torch.allclose(Y2[:kout], want, rtol=1e-12, atol=1e-12) and torch.equal(Y2[kout:], Y[kout:])
        # no basis at all (first guess block)
        Y = torch.randn((kr, n), dtype=torch.float64, device="cuda", generator=g)
        T = np.triu(np.random.default_rng(3).normal(size=(kr, kr)))
        d, g2 = k.project(None, Y, np.zeros((0, kr)), T, Y)
        assert d.shape == (0, kr) and np.allclose(g2, (Y @ Y.T).cpu().numpy(), **tol)


@pytest.mark.parametrize("nroot", [1, 3, 4])
def test_solver_matches_eigh_on_dense_matrix(torch_cuda, nroot):
    torch = torch_cuda
    from forte2_b200.davidson import DavidsonLiuSolver
    n = 600
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
    # restart from the converged vectors: one iteration
    lam2, _ = s.solve()
    assert np.max(np.abs(lam2 - want)) < 1e-11 and s.iter <= 1


def test_solver_argument_errors(torch_cuda):
    from forte2_b200.davidson import DavidsonLiuParams, DavidsonLiuSolver
    with pytest.raises(ValueError, match="smaller than 1"):
        DavidsonLiuSolver(size=0, nroot=1)
    with pytest.raises(ValueError, match="zero roots"):
        DavidsonLiuSolver(size=10, nroot=0)
    with pytest.raises(ValueError, match="collapse_per_root"):
        DavidsonLiuParams(collapse_per_root=0)
    with pytest.raises(ValueError, match="basis_per_root"):
        DavidsonLiuParams(collapse_per_root=2, basis_per_root=2)
    s = DavidsonLiuSolver(size=10, nroot=1)
    with pytest.raises(ValueError, match="Sigma builder"):
        s.solve()
    s.add_sigma_builder(lambda b, sg: None)
    with pytest.raises(ValueError, match="diagonal"):
        s.solve()
    with pytest.raises(ValueError, match="h_diag must be shape"):
        s.add_h_diag(np.zeros(3))
    with pytest.raises(ValueError, match="guesses must be shape"):
        s.add_guesses(np.zeros((3, 1)))


def test_project_out(torch_cuda):
    """davidsonliu.py:147-182: with the ground state projected out the solver finds the next root."""
    torch = torch_cuda
    from forte2_b200.davidson import DavidsonLiuSolver
    n = 300
    rng = np.random.default_rng(11)
    A = rng.normal(size=(n, n))
    A = 0.02 * (A + A.T) + np.diag(np.linspace(0.0, 10.0, n))
    w, v = np.linalg.eigh(A)
    Ad = torch.from_numpy(A).cuda()
    s = DavidsonLiuSolver(size=n, nroot=1)
    s.add_h_diag(np.diag(A).copy())
    s.add_sigma_builder(lambda Bb, Sb: Sb.copy_(Bb @ Ad))
    s.add_project_out([v[:, 0]])
    g = np.zeros((n, 2))
    g[0, 0] = g[1, 1] = 1.0
    s.add_guesses(g)
    lam, _ = s.solve()
    assert abs(lam[0] - w[1]) < 1e-10


ROOT_FIXTURES = sorted(f[len("roots_"):-4] for f in os.listdir(GOLD) if f.startswith("roots_"))


@pytest.mark.parametrize("fixture", ROOT_FIXTURES)
@pytest.mark.parametrize("nroot", [1, 3])
def test_ci_roots_match_reference(torch_cuda, fixture, nroot):
    import forte2_b200 as f2
    from forte2_b200.ci_solver import CISolver
    from forte2_b200.davidson import DavidsonLiuParams
    name, variant = fixture.rsplit("_", 1)
    space = SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, variant)
    g = np.load(os.path.join(GOLD, f"roots_{fixture}.npz"))
    strings = f2.CIStrings(*space)
    # the determinant-basis solver with the plain determinant-energy preconditioner has a slow tail on
    # some of these spaces (the reference needs up to 52 iterations, see iters_nroot* in the fixture)
    solver = CISolver(strings, E, H, V, nroot=nroot, log_level=0,
                      davidson_liu_params=DavidsonLiuParams(maxiter=300)).run()
    assert solver.converged
    want = g[f"evals_nroot{nroot}"]
    assert np.max(np.abs(solver.evals - want)) <= ROOT_TOL, (solver.evals, want)
    if "exact" in g.files:
        assert np.max(np.abs(solver.evals - g["exact"][:nroot])) <= ROOT_TOL
    # eigenvectors: orthonormal, and H v = e v through one more sigma build
    torch = torch_cuda
    v = solver.evecs
    assert torch.allclose(v @ v.T, torch.eye(nroot, dtype=torch.float64, device=v.device), atol=1e-10)
    sg = torch.empty_like(v[0])
    solver.ci_sigma_builder.Hamiltonian(v[0].contiguous(), sg)
    assert float((sg - solver.evals[0] * v[0]).norm()) < 1e-5


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


@pytest.mark.parametrize("name", ["cas66", "cas75_odd", "gas343_54", "cas66_d2h"])
def test_spin_averaged_diagonal_and_batched_hamiltonian(torch_cuda, name):
    """f2_sigma_hdiag_spin_averaged = determinant energies averaged over each orbital configuration (the S^2-invariant
    preconditioner of determinant-basis solves); f2_sigma_apply_batch = the vector-by-vector builds"""
    import forte2_b200 as f2
    torch = torch_cuda
    space = SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    e = b.form_Hdiag_det().cpu().numpy()
    got = b.form_Hdiag_det(spin_averaged=True).cpu().numpy()
    conf = {}
    for k, (wa, wb) in enumerate(strings.make_determinants()):
        conf.setdefault((int(wa & wb), int(wa ^ wb)), []).append(k)
    want = e.copy()
    for members in conf.values():
        want[members] = e[members].mean()
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    C = torch.from_numpy(np.random.default_rng(2).normal(size=(3, strings.ndet))).cuda()
    pad = torch.full((3, strings.ndet + 5), float("nan"), dtype=torch.float64, device="cuda")
    Cp, Sp = pad[:, :strings.ndet], torch.full_like(pad, float("nan"))[:, :strings.ndet]
    Cp.copy_(C)
    S1 = torch.empty_like(C)
    for i in range(3):
        b.Hamiltonian(C[i], S1[i])
    S2 = torch.full_like(C, float("nan"))
    b.Hamiltonian(C, S2)
    b.Hamiltonian(Cp, Sp)  # rows with a stride longer than ndet
    scale = float(S1.abs().max())
    assert float((S1 - S2).abs().max()) <= 1e-13 * scale and float((S1 - Sp).abs().max()) <= 1e-13 * scale


@pytest.mark.parametrize("n,multiplicity,nroot", [(8, 1, 3), (8, 3, 2), (12, 1, 3)])
def test_spin_controlled_determinant_roots_match_the_csf_path(torch_cuda, n, multiplicity, nroot):
    """Determinant-basis solve with spin control (spin-pure guesses from S^2 in the spin-complete guess space,
    wrong-spin guess combinations projected out, configuration-averaged preconditioner: forte2/sci/sci.py:405-455 and
    SURVEY.md 7.4) against the spin-adapted solve forte2 itself would run (CSF basis, forte2/ci/ci.py:214-235): same
    roots to 1e-10 Eh, and the determinant vectors lie in the CSF space of that multiplicity."""
    import forte2_b200 as f2
    from forte2_b200.ci_solver import CISolver
    torch = torch_cuda
    space = (n // 2, n // 2, 0, [[0] * n], [], [])
    E, H, V = synthetic.integrals(n, "D")
    strings = f2.CIStrings(*space)
    adapter = f2.CISpinAdapter(multiplicity - 1, 0, n)
    adapter.prepare_couplings(strings)
    csf = CISolver(strings, E, H, V, nroot=nroot, log_level=0, spin_adapter=adapter).run()
    det = CISolver(strings, E, H, V, nroot=nroot, log_level=0, multiplicity=multiplicity).run()
    assert csf.converged and det.converged
    assert np.max(np.abs(np.asarray(det.E) - np.asarray(csf.E))) <= ROOT_TOL
    back = torch.empty(adapter.ncsf, dtype=torch.float64, device="cuda")
    for i in range(nroot):
        v = det.evecs[i].contiguous()
        adapter.det_C_to_csf_C(v, back)
        assert abs(float(back.norm()) - 1.0) <= 1e-9  # nothing of the vector lies outside the spin sector
    # without spin control the same solve finds the lowest roots of ANY spin
    free = CISolver(strings, E, H, V, nroot=nroot, log_level=0).run()
    assert np.all(np.asarray(free.E) <= np.asarray(csf.E) + 1e-9)


@pytest.mark.parametrize("name", ["cas66", "gas343_54"])
def test_root_rdms_and_state_average(torch_cuda, name):
    """make_1rdm / make_2rdm of the solved roots (forte2/ci/ci.py:476-515) and their state average
    (forte2/base_classes/ci_base.py:104-133): the energy functional E0 + h.G1 + 1/2 V.G2 of root j is its eigenvalue
    (ci.py:384-392), of the weighted average the weighted energy; <i| a+_p a_q |j> of different roots is traceless"""
    import forte2_b200 as f2
    from forte2_b200.ci_solver import CISolver, make_average_rdms
    space = SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "D")
    strings = f2.CIStrings(*space)
    s = CISolver(strings, E, H, V, nroot=3, log_level=0).run()
    assert s.converged

    def functional(g1, g2):
        return E + np.einsum("pq,pq", H, g1) + 0.5 * np.einsum("pqrs,pqrs", V, g2)

    for j in range(3):
        assert abs(functional(s.make_1rdm(j), s.make_2rdm(j)) - s.E[j]) <= 1e-9
        assert abs(np.trace(s.make_1rdm(j)) - (space[0] + space[1])) <= 1e-10
    assert abs(np.trace(s.make_1rdm(0, 1))) <= 1e-9
    a, b = s.make_sd_1rdm(2)
    assert abs(np.trace(a) - space[0]) <= 1e-10 and abs(np.trace(b) - space[1]) <= 1e-10
    w = [[0.5, 0.3, 0.2]]
    g1, g2 = make_average_rdms([s], w)
    assert abs(functional(g1, g2) - float(np.dot(w[0], s.E))) <= 1e-9
    l2 = s.make_2cumulant(0)
    g1_0 = s.make_1rdm(0)
    assert np.allclose(l2, s.make_2rdm(0) - np.einsum("pr,qs->pqrs", g1_0, g1_0) + 0.5 * np.einsum("ps,qr->pqrs", g1_0, g1_0))
