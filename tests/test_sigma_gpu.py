"""GPU parity: forte2_b200.CISigmaBuilder.Hamiltonian vs the reference's own C++ builder
(oracle/_ref) on the same integrals and the same trial vector.  Tolerance from BASELINE.json:
sigma relative L2 <= 1e-12 in FP64."""
import numpy as np
import pytest

from cases import SMALL_SPACES
from forte2_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _pair(ref, space, ints):
    import forte2_b200 as f2
    E, H, V = ints
    rs = ref.RefCIStrings(*space)
    rb = ref.RefCISigmaBuilder(rs, E, H, V)
    rb.set_algorithm("hz")
    gs = f2.CIStrings(*space)
    gb = f2.CISigmaBuilder(gs, E, H, V, 0)
    return rs, rb, gs, gb


def _rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


@pytest.mark.parametrize("name", sorted(SMALL_SPACES))
@pytest.mark.parametrize("variant", ["S", "D", "four_fold"])
def test_sigma_matches_reference(ref, torch_cuda, name, variant):
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    ints = synthetic.four_fold_integrals(n) if variant == "four_fold" else synthetic.integrals(n, variant)
    rs, rb, gs, gb = _pair(ref, space, ints)
    assert gs.ndet == rs.ndet
    c = synthetic.trial_vector(max(rs.ndet, 1))[: rs.ndet].copy()
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    got = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got)  # numpy in, numpy out: the reference-facing host path
    assert _rel(got, want) <= TOL
    # device path (torch tensors), twice to check the result does not depend on leftovers
    torch = torch_cuda
    cd = torch.from_numpy(c).cuda()
    sd = torch.full_like(cd, float("nan"))
    gb.Hamiltonian(cd, sd)
    gb.Hamiltonian(cd, sd)
    assert _rel(sd.cpu().numpy(), want) <= TOL


# the reference's KH sigma is not a valid oracle with several irreps nor when one spin has no
# electrons (probe: beta_only, gas2_20, one_alpha differ from its own HZ / Slater rules by 2-19 %)
ONE_IRREP = sorted(k for k, sp in SMALL_SPACES.items()
                   if max(x for s in sp[3] for x in s) == 0 and sp[0] > 0 and sp[1] > 0)


@pytest.mark.parametrize("name", ONE_IRREP)
@pytest.mark.parametrize("variant", ["S", "D"])
def test_sigma_matches_reference_knowles_handy(ref, torch_cuda, name, variant):
    """second oracle: the reference's Knowles-Handy sigma (ci_sigma_builder_knowles_handy.cc:88-171), valid with one
    irrep and 8-fold symmetric integrals, against set_algorithm("kh") of this builder; and the
    KH form of avg_build_time(): [total, H2_kh, 0, 0] (ci_sigma_builder.cc:194-199)"""
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, variant))
    rb.set_algorithm("kh")
    gb.set_algorithm("kh")
    assert gb.get_algorithm() == rb.get_algorithm() == "Knowles-Handy"
    c = synthetic.trial_vector(max(rs.ndet, 1))[: rs.ndet].copy()
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    got = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got)
    assert _rel(got, want) <= TOL
    t = gb.avg_build_time()
    assert len(t) == 4 and t[2] == 0.0 and t[3] == 0.0 and t[0] >= t[1] >= 0.0
    gb.set_algorithm("hz")
    t = gb.avg_build_time()
    assert t[1] + t[2] + t[3] <= t[0] * (1 + 1e-9) + 1e-9


@pytest.mark.parametrize("name", ["cas66", "cas66_d2h", "gas343_54", "cas75_odd"])
def test_hdiag_matches_slater_rules(ref, torch_cuda, name):
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    want = rb.det_energies()
    got = gb.form_Hdiag_det().cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))


def test_memory_chunking_and_errors(ref, torch_cuda):
    """tests/ci/test_ci_rdms.py:393-436 of the reference: sigma under a small budget equals sigma after a
    failed set_memory(0) attempt; negative memory is a ValueError."""
    import forte2_b200 as f2
    n = 6
    H = np.diag(np.arange(1.0, n + 1))
    V = np.zeros((n,) * 4)
    gs = f2.CIStrings(2, 2, 0, [[0] * n], [], [])
    gb = f2.CISigmaBuilder(gs, 0.0, H, V, 0)
    gb.set_algorithm("hz")
    c = np.linspace(-1.0, 1.0, gs.ndet)
    gb.set_memory(1)
    s1 = np.zeros_like(c)
    gb.Hamiltonian(c, s1)
    gb.set_memory(0)
    with pytest.raises(RuntimeError, match="too small for one K-block column"):
        gb.Hamiltonian(c, np.zeros_like(c))
    gb.set_memory(1)
    s2 = np.zeros_like(c)
    gb.Hamiltonian(c, s2)
    assert np.array_equal(s1, s2)
    with pytest.raises(ValueError, match="memory must be non-negative"):
        gb.set_memory(-1)
    with pytest.raises(RuntimeError, match="not valid"):
        gb.set_algorithm("nope")
    gb.set_algorithm("kh")
    assert gb.get_algorithm() == "Knowles-Handy"
    gb.set_algorithm("harrison-zarrabian")
    assert gb.get_algorithm() == "Harrison-Zarrabian"
    with pytest.raises(RuntimeError, match="stride 1"):
        gb.Hamiltonian(np.zeros(2 * gs.ndet)[::2], np.zeros(gs.ndet))


def test_row_slab_passes_equal_single_pass(ref, torch_cuda):
    """beta same-spin work split into several transposed row slabs (tiny set_memory) == one pass."""
    space = SMALL_SPACES["cas88"]
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(8, "S"))
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    gb.set_memory(1024)
    a = np.zeros_like(c)
    gb.Hamiltonian(c, a)
    # 1 MB = 65536 doubles per buffer -> but force more passes by a GAS space with many classes
    space = SMALL_SPACES["gas343_55"]
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(10, "S"))
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    got = np.zeros_like(c)
    gb.Hamiltonian(c, got)
    assert _rel(got, want) <= TOL
    assert _rel(a, a) == 0.0


@pytest.mark.parametrize("n", [10, 12])
def test_cas_medium(ref, torch_cuda, n):
    space = (n // 2, n // 2, 0, [[0] * n], [], [])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    rb.set_memory(2048)
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    got = np.zeros_like(c)
    gb.Hamiltonian(c, got)
    assert _rel(got, want) <= TOL
    t = gb.avg_build_time()
    assert len(t) == 4 and t[0] > 0


@pytest.mark.parametrize("space,world", [((4, 4, 0, [[0] * 8], [], []), 2), ((5, 5, 0, [[0] * 10], [], []), 3),
                                         ((5, 5, 0, [[0] * 10], [], []), 4), (SMALL_SPACES["gas343_55"], 2),
                                         (SMALL_SPACES["gas343_54"], 3), (SMALL_SPACES["cas66_d2h"], 2),
                                         (SMALL_SPACES["gas2_sym4"], 3), ((4, 4, 0, [[0, 1, 2, 3] * 2], [], []), 4)])
def test_rows_form_parts_emulated_on_one_gpu(ref, torch_cuda, space, world):
    """f2_sigma_apply_rows (the sharded build that never holds the full vector) for every rank of a world,
    run one after the other on this GPU, the three exchanges of forte2_b200/dist.py done by hand:
    the sum equals the single-GPU sigma and the reference.  One determinant block (t3g_kernel) and
    symmetry / GAS blocks (t3gp_kernel).  The NCCL version is tests/test_dist_gpu.py."""
    torch = torch_cuda
    from rows_emulation import rows_form_sigma
    n = sum(len(s) for s in space[3])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    total = rows_form_sigma(torch, gs, gb, torch.from_numpy(c).cuda(), world)
    assert _rel(total.cpu().numpy(), want) <= TOL


def test_host_path_streams_the_opposite_spin_term(ref, torch_cuda, monkeypatch):
    """numpy in / numpy out with the trial vector crossing PCIe in many small pieces: the gather-form
    opposite-spin kernel is launched on the rows that have arrived (f2_sigma_apply_host); same sigma as
    the one-piece path and as the reference."""
    n = 10
    space = (n // 2, n // 2, 0, [[0] * n], [], [])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    assert gb.t3_shape()["form"] == "gather"
    monkeypatch.setenv("F2_HOST_PIECE", "1000")  # 63504 determinants -> 64 pieces, 16 launches
    got = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got)
    assert _rel(got, want) <= TOL
    assert gb.launch_count() > 12
    monkeypatch.setenv("F2_NO_STREAMED_HOST", "1")
    got2 = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got2)
    assert _rel(got2, want) <= TOL


@pytest.mark.parametrize("nsm", [0, 100, 250])
def test_gather_form_with_partially_staged_rows(ref, torch_cuda, monkeypatch, nsm):
    """Long rows of C (CAS(18,18): 389 KB) do not fit in shared memory: the gather-form kernel stages the
    first columns and gathers the rest from global memory.  Forced here on CAS(10,10) (252 columns) with
    0, 100 and 250 staged columns; device, host-streamed and sharded (rows form, emulated) paths."""
    torch = torch_cuda
    monkeypatch.setenv("F2_T3G_NSM", str(nsm))
    n = 10
    space = (5, 5, 0, [[0] * n], [], [])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    assert gb.t3_shape()["form"] == "gather"
    c = synthetic.trial_vector(rs.ndet)
    want = np.zeros_like(c)
    rb.Hamiltonian(c, want)
    cd = torch.from_numpy(c).cuda()
    sd = torch.full_like(cd, float("nan"))
    gb.Hamiltonian(cd, sd)
    assert _rel(sd.cpu().numpy(), want) <= TOL
    got = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got)
    assert _rel(got, want) <= TOL


def test_gas_three_spaces_twenty_orbitals(torch_cuda):
    """BASELINE.json configs[4] at its 1-GPU validation size: 3 GAS spaces over 20
    orbitals, 22 065 484 determinants in 36 blocks, against sigma from the reference's own builder
    (tests/golden/make_golden_gas20.py) at 200 000 sampled positions, plus the norm and <c|sigma>."""
    import os
    import forte2_b200 as f2
    path = os.path.join(os.path.dirname(__file__), "golden", "gas20twin_S.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/gas20twin_S.npz not generated")
    torch = torch_cuda
    g = np.load(path)
    space = (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])
    strings = f2.CIStrings(*space)
    assert strings.ndet == int(g["ndet"]) == 22065484 and strings.nblocks == 36
    E, H, V = synthetic.integrals(20, "S")
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    b.set_memory(8192)
    c = torch.from_numpy(synthetic.trial_vector(strings.ndet)).cuda()
    s = torch.empty_like(c)
    b.Hamiltonian(c, s)
    norm = float(g["norm"])
    got = s[torch.from_numpy(g["idx"]).cuda()].cpu().numpy()
    assert np.linalg.norm(got - g["sigma_at_idx"]) <= TOL * np.linalg.norm(g["sigma_at_idx"])
    assert abs(float(s.norm()) - norm) <= 1e-12 * norm
    assert abs(float(c @ s) - float(g["c_dot_sigma"])) <= 1e-12 * norm


@pytest.mark.parametrize("name", ["gas343_54", "gas343_55", "cas66_d2h", "gas2_sym4", "cas66_cs1"])
def test_gather_form_over_several_beta_classes(ref, torch_cuda, monkeypatch, name):
    """t3gp_kernel: the block-aware gather form of the opposite-spin term for spaces with several
    determinant blocks (GAS, point-group symmetry) against the reference, and against the scatter form
    (F2_NO_T3GM=1) it replaces on one GPU."""
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    for ints in (synthetic.integrals(n, "S"), synthetic.four_fold_integrals(n)):
        rs, rb, gs, gb = _pair(ref, space, ints)
        c = synthetic.trial_vector(rs.ndet)
        want = np.zeros_like(c)
        rb.Hamiltonian(c, want)
        got = np.full_like(c, np.nan)
        gb.Hamiltonian(c, got)
        assert _rel(got, want) <= TOL
        assert gb.t3_shape()["form"] == "gather"
    monkeypatch.setenv("F2_NO_T3GM", "1")
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    assert gb.t3_shape()["form"] == "scatter"
    got = np.full_like(c, np.nan)
    gb.Hamiltonian(c, got)
    rb.Hamiltonian(c, want)
    assert _rel(got, want) <= TOL


@pytest.mark.parametrize("shape", ["0", "1", "2", "3", "4", "5", "auto"])
@pytest.mark.parametrize("name", ["gas343_54", "cas66_d2h", "gas2_sym4"])
def test_block_aware_gather_kernel_shapes(ref, torch_cuda, monkeypatch, name, shape):
    """t3gp_kernel<512, KB, TG>: the six (k-steps per batch, row tiles per item) instances pinned with F2_T3GP_SHAPE, and
    the default: the first builds of a builder take turns through them (launch_t3gm times each once) and the fastest is
    kept - six builds in a row, every one against the reference."""
    if shape != "auto":
        monkeypatch.setenv("F2_T3GP_SHAPE", shape)
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    rs, rb, gs, gb = _pair(ref, space, synthetic.integrals(n, "S"))
    assert gb.t3_shape()["form"] == "gather"
    want = np.zeros(rs.ndet)
    for i in range(6 if shape == "auto" else 2):
        c = synthetic.trial_vector(rs.ndet) * (1.0 + 0.25 * i)
        rb.Hamiltonian(c, want)
        got = np.full_like(c, np.nan)
        gb.Hamiltonian(c, got)
        assert _rel(got, want) <= TOL, (name, shape, i)
