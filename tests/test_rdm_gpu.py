"""1- and 2-RDMs on the GPU (forte2_b200/csrc/rdm_kernels.cuh, CISigmaBuilder.a_1rdm ... sf_2cumulant) against the
definitional restatement oracle/rdm_port.py (pinned to the reference's compute_*rdm in tests/test_oracle.py) and,
where oracle/_ref is present, the reference itself; transition pairs (left != right), CAS / GAS / symmetry spaces;
and the energy functional E0 + h.g1 + 1/2 V.g2 against <c|sigma> of the device sigma build.

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


@pytest.mark.parametrize("name", ["cas66", "gas3_53", "cas66_d2h", "gas2_sym4", "cas75_odd", "beta_only"])
def test_rdms_match_the_definitions(name):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, name], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "RDM-OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])
