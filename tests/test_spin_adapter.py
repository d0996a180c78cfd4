"""CISpinAdapter (forte2_b200/ci_spin_adapter.py, csrc/spin_adapter.cu) against the reference's
forte2::CISpinAdapter: the coupling table bit-exact (CPU, golden fixtures of tests/golden/make_golden_csf.py
and, when oracle/_ref is built, the reference itself), the two transforms and the CSF preconditioner on
the GPU to 1e-14."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
import forte2_b200 as f2
from forte2_b200 import synthetic

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted((f[4:-4].rsplit("_S", 1)[0], int(f[4:-4].rsplit("_S", 1)[1])) for f in os.listdir(GOLD)
               if f.startswith("csf_"))


def _adapter(name, twoS):
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    a = f2.CISpinAdapter(twoS, space[0] - space[1], strings.norb)
    a.prepare_couplings(strings)
    return strings, a


@pytest.mark.parametrize("name,twoS", CASES)
def test_coupling_table_bit_exact_vs_golden(name, twoS):
    g = np.load(os.path.join(GOLD, f"csf_{name}_S{twoS}.npz"))
    strings, a = _adapter(name, twoS)
    bounds, det, coeff = a.table()
    assert a.ncsf == len(g["bounds"]) - 1 and a.ndet == strings.ndet and a.ncoupling == len(g["det"])
    assert np.array_equal(bounds, g["bounds"])
    assert np.array_equal(det, g["det"])
    assert np.array_equal(coeff, g["coeff"])  # same Clebsch-Gordan products, same phases: identical doubles
    ranges = a.conf_to_csfs_range()
    assert len(ranges) == a.nconf and (not ranges or (ranges[0][0] == 0 and ranges[-1][1] == a.ncsf))
    # every CSF is normalised and CSFs are orthogonal (the table is an isometry)
    dense = np.zeros((a.ncsf, strings.ndet))
    for i in range(a.ncsf):
        for d, c in a.csf(i):
            dense[i, d] += c
    assert np.allclose(dense @ dense.T, np.eye(a.ncsf), atol=1e-13)


@pytest.mark.parametrize("name", ["cas66", "cas75_odd", "gas343_54", "cas66_cs1", "one_beta_three_alpha"])
def test_coupling_table_vs_reference_all_spins(ref, name):
    space = SMALL_SPACES[name]
    twoMs = space[0] - space[1]
    rs = ref.RefCIStrings(*space)
    for twoS in range(abs(twoMs), abs(twoMs) + 7, 2):
        ra = ref.RefCISpinAdapter(rs, twoS, twoMs)
        strings, a = _adapter(name, twoS)
        rb, rd, rc = ra.table()
        gb, gd, gc = a.table()
        assert a.ncsf == ra.ncsf
        assert np.array_equal(gb, rb) and np.array_equal(gd, rd) and np.array_equal(gc, rc)


@pytest.mark.parametrize("name,twoS", [("cas66", 0), ("cas66", 2), ("cas75_odd", 1), ("gas343_54", 1),
                                       ("cas66_d2h", 0), ("gas2_sym4", 1), ("one_each", 0), ("cas88", 0)])
def test_exact_csf_preconditioner_vs_reference(ref, name, twoS):
    """form_Hdiag_csf(..., spin_adapt_full_preconditioner=True): <I|H|I> of every CSF, host Slater rules
    (forte2/ci/ci_sigma_builder.cc:326-332), against the reference's own energy_csf."""
    from forte2_b200.ci_sigma_builder import exact_hdiag_csf
    space = SMALL_SPACES[name]
    E, H, V = synthetic.integrals(sum(len(s) for s in space[3]), "S")
    rs = ref.RefCIStrings(*space)
    rb = ref.RefCISigmaBuilder(rs, E, H, V)
    ra = ref.RefCISpinAdapter(rs, twoS, space[0] - space[1])
    want = ra.hdiag_csf(rb, full=True)
    strings, a = _adapter(name, twoS)
    got = exact_hdiag_csf(strings, E, H, V, a)
    assert got.shape == want.shape and a.ncsf > 0
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    # it differs from the default (diagonal-only) form wherever a CSF has more than one determinant
    approx = ra.hdiag_csf(rb, full=False)
    if name == "cas66":
        assert np.max(np.abs(approx - want)) > 1e-6


@pytest.mark.parametrize("name,twoS", [("cas66", 0), ("cas75_odd", 1), ("gas343_54", 1), ("cas66_d2h", 0),
                                       ("gas2_sym4", 1)])
def test_csf_hamiltonian_and_elements_vs_reference(ref, name, twoS):
    """form_H_csf / slater_rules_csf / energy_csf (host helpers of the guess and of exact diagonalisation,
    forte2/ci/ci_sigma_builder.cc:278-315,372-392) against the reference's own."""
    from forte2_b200.ci_sigma_builder import csf_hamiltonian, csf_matrix_element
    space = SMALL_SPACES[name]
    E, H, V = synthetic.integrals(sum(len(s) for s in space[3]), "S")
    rs = ref.RefCIStrings(*space)
    rb = ref.RefCISigmaBuilder(rs, E, H, V)
    ra = ref.RefCISpinAdapter(rs, twoS, space[0] - space[1])
    want = ra.form_H_csf(rb)
    strings, a = _adapter(name, twoS)
    got = csf_hamiltonian(strings, E, H, V, a)
    scale = max(1.0, np.max(np.abs(want)))
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-12 * scale
    rng = np.random.default_rng(7)
    for I, J in [(0, 0), (a.ncsf - 1, 0)] + [tuple(rng.integers(0, a.ncsf, 2)) for _ in range(6)]:
        assert abs(csf_matrix_element(strings, E, H, V, a, I, J) - ra.slater_rules_csf(rb, I, J)) <= 1e-12 * scale
    # exact diagonalisation (forte2/ci/ci.py:305-318) of that matrix on the Davidson integrals: the committed
    # golden roots of the reference's CSF solve (tests/golden/make_golden_roots_csf.py)
    path = os.path.join(GOLD, f"rootscsf_{name}_S{twoS}.npz")
    if os.path.exists(path):
        g = np.load(path)
        ED, HD, VD = synthetic.integrals(strings.norb, "D")
        roots = np.linalg.eigvalsh(csf_hamiltonian(strings, ED, HD, VD, a))[: g["exact"].size]
        assert np.max(np.abs(roots - g["exact"][: roots.size])) <= 1e-10


def test_errors():
    strings = f2.CIStrings(*SMALL_SPACES["cas66"])
    a = f2.CISpinAdapter(0, 2, 6)
    with pytest.raises(ValueError, match="twoMs"):
        a.prepare_couplings(strings)
    a = f2.CISpinAdapter(0, 0, 7)
    with pytest.raises(ValueError, match="orbitals"):
        a.prepare_couplings(strings)
    a = f2.CISpinAdapter(0, 0, 6)
    with pytest.raises(RuntimeError, match="prepare_couplings"):
        a.ncsf
    with pytest.raises(TypeError):
        a.prepare_couplings(strings.make_determinants())
    a.prepare_couplings(strings.make_determinants(), strings)
    assert a.ncsf == 175


@pytest.mark.gpu
@pytest.mark.parametrize("name,twoS", CASES)
def test_transforms_and_preconditioner_on_gpu(name, twoS):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    g = np.load(os.path.join(GOLD, f"csf_{name}_S{twoS}.npz"))
    strings, a = _adapter(name, twoS)
    # numpy-facing calls (what forte2's Davidson closure issues, forte2/ci/ci.py:228-233)
    to_det = np.full(strings.ndet, np.nan)
    a.csf_C_to_det_C(g["c_csf"], to_det)
    assert np.max(np.abs(to_det - g["to_det"]), initial=0.0) <= 1e-14
    to_csf = np.full(a.ncsf, np.nan)
    a.det_C_to_csf_C(g["c_det"], to_csf)
    assert np.max(np.abs(to_csf - g["to_csf"]), initial=0.0) <= 1e-14
    # device tensors
    d_det = torch.full((strings.ndet,), float("nan"), dtype=torch.float64, device="cuda")
    a.csf_C_to_det_C(torch.from_numpy(g["c_csf"]).cuda(), d_det)
    assert np.max(np.abs(d_det.cpu().numpy() - g["to_det"]), initial=0.0) <= 1e-14
    d_csf = torch.full((a.ncsf,), float("nan"), dtype=torch.float64, device="cuda")
    a.det_C_to_csf_C(torch.from_numpy(g["c_det"]).cuda(), d_csf)
    assert np.max(np.abs(d_csf.cpu().numpy() - g["to_csf"]), initial=0.0) <= 1e-14
    # CSF preconditioner
    E, H, V = synthetic.integrals(strings.norb, "S")
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    hd = b.form_Hdiag_csf(None, a).cpu().numpy()
    assert np.max(np.abs(hd - g["hdiag_csf"]), initial=0.0) <= 1e-12 * max(1.0, np.max(np.abs(g["hdiag_csf"]), initial=0.0))
    with pytest.raises(ValueError):
        a.csf_C_to_det_C(torch.zeros(a.ncsf + 1, dtype=torch.float64, device="cuda"), d_det)


ROOT_CASES = sorted((f[9:-4].rsplit("_S", 1)[0], int(f[9:-4].rsplit("_S", 1)[1])) for f in os.listdir(GOLD)
                    if f.startswith("rootscsf_"))


@pytest.mark.gpu
@pytest.mark.parametrize("name,twoS", ROOT_CASES)
@pytest.mark.parametrize("nroot", [1, 3])
def test_spin_adapted_roots_match_reference(name, twoS, nroot):
    """CISolver in the CSF basis (device adapter + device builder + device Davidson-Liu) against the roots
    the reference's adapter + builder + DavidsonLiuSolver give on the same integrals
    (tests/golden/make_golden_roots_csf.py).  BASELINE.json: converged roots <= 1e-10 Eh."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from forte2_b200.ci_solver import CISolver
    from forte2_b200.davidson import DavidsonLiuParams
    g = np.load(os.path.join(GOLD, f"rootscsf_{name}_S{twoS}.npz"))
    strings, a = _adapter(name, twoS)
    assert a.ncsf == int(g["ncsf"])
    E, H, V = synthetic.integrals(strings.norb, "D")
    solver = CISolver(strings, E, H, V, nroot=nroot, davidson_liu_params=DavidsonLiuParams(r_tol=1e-7), log_level=0,
                      spin_adapter=a).run()
    assert solver.converged
    assert np.max(np.abs(np.asarray(solver.E) - g[f"evals_nroot{nroot}"])) <= 1e-10
    assert np.max(np.abs(np.asarray(solver.E) - g["exact"][:nroot])) <= 1e-10
    # the roots are eigenvectors of H in the determinant basis too: |H v - E v| small
    v = solver.evecs_det()
    hv = torch.empty_like(v[0])
    solver.ci_sigma_builder.Hamiltonian(v[0].contiguous(), hv)
    assert float((hv - float(solver.E[0]) * v[0]).norm()) <= 1e-5


@pytest.fixture
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


@pytest.mark.gpu
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



@pytest.mark.gpu
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
