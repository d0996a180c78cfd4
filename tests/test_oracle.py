"""The oracle checks itself: the C/Python restatement (oracle/sigma_port.c, oracle/strings_port.py)
against the committed golden vectors generated from the reference, and against the live reference
library (oracle/_ref) where it has been built."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
import forte2_b200 as f2
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


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas21_4", "one_each", "beta_only"])
def test_rdm_port_matches_live_reference(ref, name):
    """oracle/rdm_port.py (RDMs from their definitions over the determinant list) against the reference's own
    compute_*rdm for a transition pair (left != right): the pinned checker of the next widening step
    (SURVEY.md section 8f, rank 2)."""
    from oracle import rdm_port
    space = SMALL_SPACES[name]
    norb = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(norb, "S")
    rs = ref.RefCIStrings(*space)
    rb = ref.RefCISigmaBuilder(rs, E, H, V)
    dets = f2.CIStrings(*space).make_determinants()
    rng = np.random.default_rng(11)
    left, right = rng.normal(size=rs.ndet), rng.normal(size=rs.ndet)
    left /= np.linalg.norm(left)
    right /= np.linalg.norm(right)
    checks = [("a_1rdm", rdm_port.s_1rdm(dets, norb, left, right, 0)),
              ("b_1rdm", rdm_port.s_1rdm(dets, norb, left, right, 1)),
              ("sf_1rdm", rdm_port.sf_1rdm(dets, norb, left, right)),
              ("ab_2rdm", rdm_port.ab_2rdm(dets, norb, left, right))]
    if norb >= 2:
        checks += [("aa_2rdm", rdm_port.ss_2rdm(dets, norb, left, right, 0)),
                   ("bb_2rdm", rdm_port.ss_2rdm(dets, norb, left, right, 1)),
                   ("sf_2rdm", rdm_port.sf_2rdm(dets, norb, left, right)),
                   ("sf_2cumulant", rdm_port.sf_2cumulant(dets, norb, left, right))]
    for kind, got in checks:
        want = rb.rdm(kind, left, right)
        assert got.shape == want.shape, kind
        assert np.max(np.abs(got - want), initial=0.0) <= 1e-13, kind
    # the energy functional closes the loop with the sigma oracle: E = E0 + h.g1 + 1/2 V.g2 (forte2/ci/ci.py:356-363)
    c = right
    sigma = np.zeros_like(c)
    rb.set_algorithm("hz")  # the reference's KH path is not P H P with several irreps
    rb.Hamiltonian(c, sigma)
    g1 = rb.rdm("sf_1rdm", c, c)
    g2 = rb.rdm("sf_2rdm", c, c)
    energy = E + np.einsum("pq,pq", H, g1) + 0.5 * np.einsum("pqrs,pqrs", V, g2)
    assert abs(energy - c @ sigma) <= 1e-11


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas21_4", "one_each", "cas75_odd",
                                  "one_beta_three_alpha", "beta_only", "sym2_11"])
def test_rdm_kernel_models_match_the_definitions(name):
    """tests/rdm_kernel_model.py (the RDM kernels of csrc/rdm_kernels.cuh restated statement by statement on the
    host, reading the device-format tables) against oracle/rdm_port.py, itself pinned to the reference: list
    semantics, unpacking and offset arithmetic of the kernels are right before they ever run on a GPU."""
    from oracle import rdm_port
    import rdm_kernel_model as km
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    norb = strings.norb
    dets = strings.make_determinants()
    rng = np.random.default_rng(3)
    left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
    t = km.Tables(strings)
    tol = 1e-13
    if space[0] >= 1:
        assert np.max(np.abs(km.rdm1a(t, left, right) - rdm_port.s_1rdm(dets, norb, left, right, 0))) <= tol
    if space[1] >= 1:
        assert np.max(np.abs(km.rdm1b(t, left, right) - rdm_port.s_1rdm(dets, norb, left, right, 1))) <= tol
    if space[0] >= 1 and space[1] >= 1:
        assert np.max(np.abs(km.rdm2ab(t, left, right) - rdm_port.ab_2rdm(dets, norb, left, right))) <= tol
    if space[0] >= 2:
        assert np.max(np.abs(km.rdm2ss(t, left, right, 0) - rdm_port.ss_2rdm(dets, norb, left, right, 0))) <= tol
    if space[1] >= 2:
        assert np.max(np.abs(km.rdm2ss(t, left, right, 1) - rdm_port.ss_2rdm(dets, norb, left, right, 1))) <= tol


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas21_4"])
def test_sf_2rdm_assembly_matches_the_definitions(name):
    """forte2_b200.ci_sigma_builder.assemble_sf_2rdm (host assembly of the spin-free 2-RDM from the spin blocks the
    kernels produce) against oracle/rdm_port.sf_2rdm"""
    from oracle import rdm_port
    from forte2_b200.ci_sigma_builder import assemble_sf_2rdm
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    norb, dets = strings.norb, strings.make_determinants()
    rng = np.random.default_rng(8)
    left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
    got = assemble_sf_2rdm(norb, rdm_port.ab_2rdm(dets, norb, left, right),
                           rdm_port.ss_2rdm(dets, norb, left, right, 0), rdm_port.ss_2rdm(dets, norb, left, right, 1))
    assert np.max(np.abs(got - rdm_port.sf_2rdm(dets, norb, left, right))) <= 1e-13


@pytest.fixture(scope="module")
def rdm_emu(tmp_path_factory):
    """the RDM kernels' own source (csrc/rdm_kernels.cuh) compiled for the host under the SIMT emulator of
    tests/simt_emu (one std::thread per CUDA thread, barriers for __syncthreads and the shuffles)"""
    import ctypes
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    out = tmp_path_factory.mktemp("simt")
    script = os.path.join(os.path.dirname(__file__), "simt_emu", "build.sh")
    r = subprocess.run(["bash", script, str(out)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return ctypes.CDLL(os.path.join(str(out), "librdm_emu.so"))


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas21_4", "beta_only"])
def test_rdm_kernel_source_under_the_simt_emulator(rdm_emu, name):
    """The CUDA text of rdm1a / rdm1b / rdm2ab / rdm2ss<0,1> executed thread by thread on the host (64 threads per
    block: two warps, so the cross-warp reduction, the shuffles, the shared-memory atomics and every barrier are
    exercised) against oracle/rdm_port.py.  Together with the host models this checks everything about these
    kernels that does not need the hardware."""
    import ctypes
    from oracle import rdm_port
    import rdm_kernel_model as km
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    norb, dets = strings.norb, strings.make_determinants()
    t = km.Tables(strings)
    rng = np.random.default_rng(4)
    left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)  # noqa: E731
    u64 = lambda a: np.ascontiguousarray(a).astype(np.uint64) if a.dtype != np.uint64 else np.ascontiguousarray(a)  # noqa: E731
    keep = dict(a_cls=u32(t.a_cls), a_loc=u32(t.a_loc), b_cls=u32(t.b_cls), b_loc=u32(t.b_loc), nA=u32(t.nA_cls),
                nB=u32(t.nB_cls), blk=np.ascontiguousarray(t.blk_off, dtype=np.int64), as_ptr=u32(t.as_ptr),
                as_ent=u64(t.as_ent), bs_ent=u64(t.bs_ent) if t.bs_ent.size else np.zeros(1, dtype=np.uint64))
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731

    def run(kind, count, spin=0):
        out = np.zeros(max(count, 1))
        h2p, h2e = u32(t.h2_ptr[spin]), u64(t.h2_ent[spin]) if t.h2_ent[spin].size else np.zeros(1, dtype=np.uint64)
        rc = rdm_emu.emu_rdm(kind, 64, ptr(left), ptr(right), ptr(out), ptr(keep["a_cls"]), ptr(keep["a_loc"]),
                             ptr(keep["b_cls"]), ptr(keep["b_loc"]), ptr(keep["nA"]), ptr(keep["nB"]), ptr(keep["blk"]),
                             ptr(keep["as_ptr"]), ptr(keep["as_ent"]), ptr(keep["bs_ent"]), ptr(h2p), ptr(h2e),
                             max(len(h2p) - 1, 0), norb, t.nclsA, t.nclsB, t.bs_width,
                             ctypes.c_int64(t.nas), ctypes.c_int64(t.nbs))
        assert rc == 0
        return out[:count]

    n2, npairs = norb * norb, norb * (norb - 1) // 2
    tol = 1e-12
    if space[0] >= 1:
        assert np.max(np.abs(run(0, n2) - rdm_port.s_1rdm(dets, norb, left, right, 0).reshape(-1))) <= tol
    if space[1] >= 1:
        assert np.max(np.abs(run(1, n2) - rdm_port.s_1rdm(dets, norb, left, right, 1).reshape(-1))) <= tol
    if space[0] >= 1 and space[1] >= 1:
        assert np.max(np.abs(run(5, n2 * n2) - rdm_port.ab_2rdm(dets, norb, left, right).reshape(-1))) <= tol
    if space[0] >= 2:
        assert np.max(np.abs(run(3, npairs * npairs, 0) - rdm_port.ss_2rdm(dets, norb, left, right, 0).reshape(-1))) <= tol
    if space[1] >= 2:
        assert np.max(np.abs(run(4, npairs * npairs, 1) - rdm_port.ss_2rdm(dets, norb, left, right, 1).reshape(-1))) <= tol


# ---- the sampled Slater-rules checker (oracle/sampled.py): what the parity tests at 10^8 - 10^9 determinants rest on
@pytest.mark.parametrize("fixture", SIGMA_FIXTURES)
def test_sampled_checker_matches_golden_sigma(fixture):
    """every sampled sigma_I rebuilt from Slater rules + the restated string layout equals the reference's sigma
    (committed goldens: CAS, GAS, point-group symmetry, na != nb; 8-fold and 4-fold symmetric integrals)"""
    from oracle import sampled
    name, variant = fixture.rsplit("_", 1) if not fixture.endswith("four_fold") else (fixture[:-10], "four_fold")
    space = SMALL_SPACES[name]
    g = np.load(os.path.join(GOLD, f"sigma_{fixture}.npz"))
    lay = sampled.layout(space)
    assert lay.ndet == len(g["c"])
    assert np.array_equal(lay.address(g["dets"][:, 0], g["dets"][:, 1]), np.arange(lay.ndet))
    wa, wb = lay.determinant(np.arange(lay.ndet))
    assert np.array_equal(wa, g["dets"][:, 0]) and np.array_equal(wb, g["dets"][:, 1])
    n = sum(len(s) for s in space[3])
    E, H, V = _ints(n, variant)
    err, ns = sampled.sampled_rel_error(space, E, H, V, g["c"], g["sigma_hz"], nsamp=400)
    assert ns > 0 and err <= 1e-13
    if lay.ndet > 20:  # and it is not vacuous
        bad = g["sigma_hz"].copy()
        bad[sampled.sample_addresses(lay.ndet, 400)[0]] += 1e-9 * np.linalg.norm(bad)
        assert sampled.sampled_rel_error(space, E, H, V, g["c"], bad, nsamp=400)[0] > 1e-10


@pytest.mark.parametrize("name", ["gas343_55", "cas66_d2h", "gas2_sym4", "cas88"])
def test_sampled_checker_matches_live_reference(ref, name):
    from oracle import sampled
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = _ints(n, "D")
    c = synthetic.trial_vector(ref.RefCIStrings(*space).ndet, seed=11)
    want, kind = checker.sigma(space, E, H, V, c)
    assert kind == "reference"
    err, _ = sampled.sampled_rel_error(space, E, H, V, c, want, nsamp=500)
    assert err <= 1e-13


def test_layout_of_large_spaces_without_enumerating_determinants():
    """sizes and block counts of BASELINE configs 3-5 from the restated layout alone"""
    from oracle import sampled
    assert sampled.layout((9, 9, 0, [[0] * 18], [], [])).ndet == 2363904400
    lay = sampled.layout((10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [8, 0, 0], [12, 16, 4]))
    assert lay.ndet == 1865504274 and sum(1 for b in lay.blocks if b[3] * b[4] > 0) == 225
    lay = sampled.layout((10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2]))
    assert lay.ndet == 22065484 and len(lay.blocks) == 36
    strings = f2.CIStrings(10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])
    adr = sampled.sample_addresses(lay.ndet, 300)
    wa, wb = lay.determinant(adr)
    for k in range(0, len(adr), 37):  # the product's own addressing agrees
        assert strings.determinant(int(adr[k])) == (int(wa[k]), int(wb[k]))
