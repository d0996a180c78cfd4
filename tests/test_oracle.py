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


def _hole_tables(strings, spin, level):
    ptr = strings.export(f"hole_ptr{level}", spin)
    ent = strings.export(f"hole_ent{level}", spin).astype(np.uint64)
    return ptr, ent


def _gather_model(strings, x, la, lb):
    """rdm_gather_kernel of csrc/rdm_kernels.cuh restated with numpy over the exported tables: the dense D matrix
    D[(opA Pb + opB), (Ka, Kb)] = s_A s_B x[Ka + A, Kb + B] (all columns at once)"""
    from math import comb
    n = strings.norb
    Pa, Pb = comb(n, la), comb(n, lb)
    ptrA, entA = _hole_tables(strings, 0, la)
    ptrB, entB = _hole_tables(strings, 1, lb)
    a_cls, a_loc = strings.export("dev_cls", 0), strings.export("dev_loc", 0)
    b_cls, b_loc = strings.export("dev_cls", 1), strings.export("dev_loc", 1)
    nB = strings.export("dev_ncls", 1)
    off = strings.export("dev_blk_off", 0).reshape(len(strings.export("dev_ncls", 0)), len(nB))
    nKa, nKb = len(ptrA) - 1, len(ptrB) - 1
    D = np.zeros((Pa * Pb, max(nKa, 0) * max(nKb, 0)))
    for Ka in range(nKa):
        for ea in entA[ptrA[Ka]:ptrA[Ka + 1]]:
            I, opA, negA = int(ea & 0xffffffff), int((ea >> 32) & 0xffff), int((ea >> 48) & 1)
            for Kb in range(nKb):
                for eb in entB[ptrB[Kb]:ptrB[Kb + 1]]:
                    J, opB, negB = int(eb & 0xffffffff), int((eb >> 32) & 0xffff), int((eb >> 48) & 1)
                    o = off[a_cls[I], b_cls[J]]
                    if o < 0:
                        continue
                    v = x[o + a_loc[I] * nB[b_cls[J]] + b_loc[J]]
                    D[opA * Pb + opB, Ka * nKb + Kb] = -v if negA ^ negB else v
    return D


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas21_4", "one_each", "cas75_odd",
                                  "one_beta_three_alpha", "beta_only", "sym2_11"])
def test_rdm_gather_gemm_formulation_matches_the_definitions(name):
    """Gamma = D_L D_R^T with D gathered through the hole lists (csrc/strings.hpp HoleLists; the formulation of
    csrc/rdm_kernels.cuh, restated over the tables the kernels read) against oracle/rdm_port.general_rdm for every
    spin block up to three-body: list contents, packing, phases and offsets are right before the kernels run"""
    from oracle import rdm_port
    space = SMALL_SPACES[name]
    strings = f2.CIStrings(*space)
    norb, dets = strings.norb, strings.make_determinants()
    rng = np.random.default_rng(3)
    left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
    levels = [(1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]
    if strings.ndet <= 400:
        levels += [(3, 0), (2, 1), (1, 2), (0, 3)]
    for la, lb in levels:
        if space[0] < la or space[1] < lb:
            continue
        got = _gather_model(strings, left, la, lb) @ _gather_model(strings, right, la, lb).T
        want = rdm_port.general_rdm(dets, norb, left, right, la, lb)
        assert got.shape == want.shape and np.max(np.abs(got - want), initial=0.0) <= 1e-12, (la, lb)


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas21_4", "cas75_odd", "one_each", "beta_only"])
def test_general_rdm_port_and_3rdm_assembly_match_live_reference(ref, name):
    """oracle/rdm_port.general_rdm (any spin block from the definition) against the reference's compute_*_3rdm
    (ci_sigma_builder_3rdm.cc) and its 1-/2-body blocks; the host assembly of the spin-free 3-RDM and 3-cumulant
    (forte2_b200.ci_sigma_builder.assemble_sf_3rdm / sf_3cumulant_from) against compute_sf_3rdm / compute_sf_3cumulant"""
    from oracle import rdm_port
    from forte2_b200.ci_sigma_builder import assemble_sf_3rdm, sf_3cumulant_from
    space = SMALL_SPACES[name]
    norb = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(norb, "S")
    rs = ref.RefCIStrings(*space)
    rb = ref.RefCISigmaBuilder(rs, E, H, V)
    dets = f2.CIStrings(*space).make_determinants()
    rng = np.random.default_rng(12)
    left, right = rng.normal(size=rs.ndet), rng.normal(size=rs.ndet)
    left /= np.linalg.norm(left)
    right /= np.linalg.norm(right)
    blocks = {}
    for kind, (la, lb) in {"a_1rdm": (1, 0), "b_1rdm": (0, 1), "aa_2rdm": (2, 0), "bb_2rdm": (0, 2), "ab_2rdm": (1, 1),
                           "aaa_3rdm": (3, 0), "aab_3rdm": (2, 1), "abb_3rdm": (1, 2), "bbb_3rdm": (0, 3)}.items():
        if norb < max(la, lb) or (la and lb and norb < 2):
            continue
        want = rb.rdm(kind, left, right)
        got = rdm_port.general_rdm(dets, norb, left, right, la, lb)
        assert got.size == want.size, kind
        assert np.max(np.abs(got.reshape(want.shape) - want), initial=0.0) <= 1e-13, kind
        blocks[kind] = got.reshape(want.shape)
    if norb >= 2:
        sss = blocks["aaa_3rdm"] + blocks["bbb_3rdm"] if norb >= 3 else None
        g3 = assemble_sf_3rdm(norb, blocks["aab_3rdm"], blocks["abb_3rdm"], sss)
        want3 = rb.rdm("sf_3rdm", left, right)
        assert np.max(np.abs(g3 - want3), initial=0.0) <= 1e-13
        l3 = sf_3cumulant_from(rb.rdm("sf_1rdm", left, right), rb.rdm("sf_2rdm", left, right), want3)
        assert np.max(np.abs(l3 - rb.rdm("sf_3cumulant", left, right)), initial=0.0) <= 1e-12
    if norb <= 4 and norb >= 2:
        assert np.max(np.abs(rdm_port.sf_3rdm(dets, norb, left, right) - rb.rdm("sf_3rdm", left, right))) <= 1e-13


TRDM_PAIRS = {
    # (left space, right space): same orbitals and electron numbers, different symmetry / GAS restrictions
    "sym_0_3": ((3, 3, 0, [[0, 1, 2, 3, 0, 1]], [], []), (3, 3, 3, [[0, 1, 2, 3, 0, 1]], [], [])),
    "sym_1_2_odd": ((3, 2, 1, [[0, 1, 0, 1, 1]], [], []), (3, 2, 0, [[0, 1, 0, 1, 1]], [], [])),
    "gas_vs_cas": ((3, 3, 0, [[0] * 3, [0] * 3], [2, 0], [6, 4]), (3, 3, 0, [[0] * 3, [0] * 3], [], [])),
    "gas_vs_gas": ((3, 2, 0, [[0] * 2, [0] * 4], [2, 0], [3, 3]), (3, 2, 0, [[0] * 2, [0] * 4], [3, 0], [4, 2])),
    "same": ((3, 3, 0, [[0] * 6], [], []), (3, 3, 0, [[0] * 6], [], [])),
}


@pytest.mark.parametrize("pair", sorted(TRDM_PAIRS))
def test_trdm_port_matches_live_reference(ref, pair):
    """oracle/rdm_port.trdm1 against the reference's compute_{a,b,sf}_1trdm (ci_sigma_builder_1trdm.cc)"""
    from oracle import rdm_port
    sl, sr = TRDM_PAIRS[pair]
    norb = sum(len(s) for s in sl[3])
    E, H, V = synthetic.integrals(norb, "S")
    bl = ref.RefCISigmaBuilder(ref.RefCIStrings(*sl), E, H, V)
    br = ref.RefCISigmaBuilder(ref.RefCIStrings(*sr), E, H, V)
    dl, dr = f2.CIStrings(*sl).make_determinants(), f2.CIStrings(*sr).make_determinants()
    rng = np.random.default_rng(5)
    left, right = rng.normal(size=len(dl)), rng.normal(size=len(dr))
    a = rdm_port.trdm1(dl, dr, norb, left, right, 0)
    b = rdm_port.trdm1(dl, dr, norb, left, right, 1)
    assert np.max(np.abs(a - bl.trdm("a_1trdm", br, left, right))) <= 1e-13
    assert np.max(np.abs(b - bl.trdm("b_1trdm", br, left, right))) <= 1e-13
    assert np.max(np.abs(a + b - bl.trdm("sf_1trdm", br, left, right))) <= 1e-13


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
