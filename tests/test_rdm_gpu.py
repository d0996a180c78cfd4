"""1-, 2- and 3-RDMs, cumulants and transition 1-RDMs on the GPU (forte2_b200/csrc/rdm_kernels.cuh,
CISigmaBuilder.a_1rdm ... sf_3cumulant, a_1trdm ... sf_1trdm) against the definitional restatement oracle/rdm_port.py
(pinned to the reference's compute_*rdm in tests/test_oracle.py) and, where oracle/_ref is present, the reference itself;
transition pairs (left != right), CAS / GAS / symmetry spaces; the energy functional E0 + h.g1 + 1/2 V.g2 against
<c|sigma> of the device sigma build; several column chunks per block; the memory-limit error of the reference.

Every case runs in its own process (a faulting kernel cannot poison the CUDA context of the suite)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
import numpy as np
root, name = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + "/tests")
import torch
import forte2_b200 as f2
from forte2_b200 import synthetic
from oracle import rdm_port, ref
from cases import SMALL_SPACES

space = SMALL_SPACES[name]
strings = f2.CIStrings(*space)
norb, dets = strings.norb, strings.make_determinants()
E, H, V = synthetic.integrals(norb, "S")
b = f2.CISigmaBuilder(strings, E, H, V, 0)
rng = np.random.default_rng(17)
left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
left /= np.linalg.norm(left)
right /= np.linalg.norm(right)
want = {"a_1rdm": rdm_port.s_1rdm(dets, norb, left, right, 0), "b_1rdm": rdm_port.s_1rdm(dets, norb, left, right, 1),
        "sf_1rdm": rdm_port.sf_1rdm(dets, norb, left, right), "ab_2rdm": rdm_port.ab_2rdm(dets, norb, left, right),
        "aa_2rdm": rdm_port.ss_2rdm(dets, norb, left, right, 0), "bb_2rdm": rdm_port.ss_2rdm(dets, norb, left, right, 1),
        "sf_2rdm": rdm_port.sf_2rdm(dets, norb, left, right),
        "sf_2cumulant": rdm_port.sf_2cumulant(dets, norb, left, right)}
rb = None
if ref.available():
    rb = ref.RefCISigmaBuilder(ref.RefCIStrings(*space), E, H, V)
worst = 0.0
for kind, w in want.items():
    got = getattr(b, kind)(left, right)
    assert got.shape == w.shape, (kind, got.shape, w.shape)
    err = float(np.max(np.abs(got - w), initial=0.0))
    if rb is not None:
        err = max(err, float(np.max(np.abs(got - rb.rdm(kind, left, right)), initial=0.0)))
    print(kind, err)
    worst = max(worst, err)
# three-body blocks, spin-free 3-RDM and 3-cumulant (ci_sigma_builder_3rdm.cc)
if strings.ndet <= 1000:
    from math import comb
    for kind, (la, lb) in {"aaa_3rdm": (3, 0), "aab_3rdm": (2, 1), "abb_3rdm": (1, 2), "bbb_3rdm": (0, 3)}.items():
        got = getattr(b, kind)(left, right)
        if norb < 3 and kind in ("aaa_3rdm", "bbb_3rdm") or norb < 2:
            assert got.size == 0, kind
            continue
        w = rdm_port.general_rdm(dets, norb, left, right, la, lb).reshape(got.shape)
        err = float(np.max(np.abs(got - w), initial=0.0))
        if rb is not None:
            err = max(err, float(np.max(np.abs(got - rb.rdm(kind, left, right)), initial=0.0)))
        print(kind, err)
        worst = max(worst, err)
    if rb is not None:
        for kind in ("sf_3rdm", "sf_3cumulant"):
            err = float(np.max(np.abs(getattr(b, kind)(left, right) - rb.rdm(kind, left, right)), initial=0.0))
            print(kind, err)
            worst = max(worst, err)
# device tensors in, same result
gl, gr = torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()
worst = max(worst, float(np.max(np.abs(b.ab_2rdm(gl, gr) - want["ab_2rdm"]), initial=0.0)))
# energy functional against the device sigma build (forte2/ci/ci.py:384-392)
c = gr
sigma = torch.empty_like(c)
b.Hamiltonian(c, sigma)
g1, g2 = b.sf_1rdm(c, c), b.sf_2rdm(c, c)
energy = E + np.einsum("pq,pq", H, g1) + 0.5 * np.einsum("pqrs,pqrs", V, g2)
gap = abs(energy - float(c @ sigma))
print("energy functional", gap)
assert worst <= 1e-12 and gap <= 1e-11, (worst, gap)
print("RDM-OK")
"""

TRDM_SCRIPT = r"""
import sys
import numpy as np
root, pair = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
sys.path.insert(0, root + "/tests")
import torch
import forte2_b200 as f2
from forte2_b200 import synthetic
from oracle import rdm_port, ref
from test_oracle import TRDM_PAIRS

sl, sr = TRDM_PAIRS[pair]
norb = sum(len(s) for s in sl[3])
E, H, V = synthetic.integrals(norb, "S")
SL, SR = f2.CIStrings(*sl), f2.CIStrings(*sr)
bl, br = f2.CISigmaBuilder(SL, E, H, V, 0), f2.CISigmaBuilder(SR, E, H, V, 0)
dl, dr = SL.make_determinants(), SR.make_determinants()
rng = np.random.default_rng(5)
left, right = rng.normal(size=SL.ndet), rng.normal(size=SR.ndet)
a, b = rdm_port.trdm1(dl, dr, norb, left, right, 0), rdm_port.trdm1(dl, dr, norb, left, right, 1)
worst = max(np.max(np.abs(bl.a_1trdm(br, left, right) - a)), np.max(np.abs(bl.b_1trdm(br, left, right) - b)),
            np.max(np.abs(bl.sf_1trdm(br, torch.from_numpy(left).cuda(), torch.from_numpy(right).cuda()) - a - b)))
if ref.available():
    rl = ref.RefCISigmaBuilder(ref.RefCIStrings(*sl), E, H, V)
    rr = ref.RefCISigmaBuilder(ref.RefCIStrings(*sr), E, H, V)
    worst = max(worst, np.max(np.abs(bl.sf_1trdm(br, left, right) - rl.trdm("sf_1trdm", rr, left, right))))
# the reference's argument checks (ci_sigma_builder_1trdm.cc:24-34)
other = f2.CISigmaBuilder(f2.CIStrings(sl[0], sl[1] - 1, 0, [[0] * norb], [], []), E, H, V, 0)
try:
    bl.a_1trdm(other, left, np.zeros(other.ci_strings.ndet))
    raise SystemExit("no error for different electron numbers")
except RuntimeError as e:
    assert "number of alpha and beta electrons" in str(e), e
print("worst", worst)
assert worst <= 1e-12, worst
print("TRDM-OK")
"""


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas75_odd", "beta_only", "cas88", "cas21_4",
                                  "one_each", "gas343_54"])
def test_rdms_match_the_definitions(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, name], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "RDM-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas75_odd"])
def test_rdms_in_several_column_chunks(name):
    """the chunk loop over the composite hole index (K blocks sized by set_memory, ci_sigma_builder_3rdm.cc:235-258):
    37 columns per chunk on small spaces"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, F2_RDM_CHUNK_COLS="37")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, name], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "RDM-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h"])
def test_rdms_through_the_d_matrix_path(name):
    """F2_RDM_GENERIC: every block through the gather + GEMM formulation (what the mixed-spin 3-RDM blocks and the spaces
    outside the envelope of rdm_gram_kernel / rdm_ab_kernel always use)"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, F2_RDM_GENERIC="1")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, name], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "RDM-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("name,nsm", [("cas66", 7), ("cas88", 0), ("cas88", 33), ("cas75_odd", 16)])
def test_ab_2rdm_long_row_variant(name, nsm):
    """F2_RDM_AB_NSM: rdm_ab_kernel<..., LONG> (rows of R longer than shared memory, CAS(18,18)): only nsm columns of the
    row are staged, the others are gathered from global memory"""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, F2_RDM_AB_NSM=str(nsm))
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, name], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "RDM-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


def test_transition_1rdms_between_two_spaces():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from test_oracle import TRDM_PAIRS
    for pair in sorted(TRDM_PAIRS):
        r = subprocess.run([sys.executable, "-c", TRDM_SCRIPT, ROOT, pair], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "TRDM-OK" in r.stdout, (pair, r.stdout[-2000:], r.stderr[-3000:])


def test_rdm_memory_limit_error():
    """a K-block column that does not fit set_memory raises the reference's message
    (acquire_local_Kblock_buffers, ci_sigma_builder.cc:70-78)"""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    strings = f2.CIStrings(3, 3, 0, [[0] * 6], [], [])
    E, H, V = synthetic.integrals(6, "S")
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    c = np.ones(strings.ndet) / np.sqrt(strings.ndet)
    b.set_memory(1)
    assert b.aab_3rdm(c, c).shape == (15, 6, 15, 6)
    b.set_memory(0)
    with pytest.raises(RuntimeError, match="too small for one K-block column"):
        b.aab_3rdm(c, c)


@pytest.mark.parametrize("name", ["cas88", "cas66_d2h", "gas343_54"])
def test_rdm_parts_add_up(name):
    """f2_sigma_rdm_part: the blocks of the parts (slices of the hole strings / input strings, one per GPU in a sharded
    run) add up to the block, for the Gram, the gather-form and the D-matrix kernels"""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    from cases import SMALL_SPACES
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    rng = np.random.default_rng(23)
    left, right = rng.normal(size=strings.ndet), rng.normal(size=strings.ndet)
    for kind in ("a_1rdm", "b_1rdm", "aa_2rdm", "bb_2rdm", "ab_2rdm", "aab_3rdm"):
        want = getattr(b, kind)(left, right)
        for nparts in (2, 3, 8):
            got = sum(getattr(b, kind)(left, right, part=(i, nparts)) for i in range(nparts))
            assert np.max(np.abs(got - want), initial=0.0) <= 1e-12 * max(1.0, np.max(np.abs(want), initial=0.0)), (kind, nparts)
