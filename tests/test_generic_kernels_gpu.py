"""The shape-generic sigma kernels (t3_generic_kernel / ss_generic_kernel, csrc/kernels.cuh): spaces whose replacement
lists are longer than the tensor-core kernels hold -- more than 128 single replacements per alpha string, more than 112
parents per (N-2)-electron string -- which the reference handles like any other space.  Forced on the small spaces
(F2_GENERIC_KERNELS) against the reference builder, and selected by shape on few-electron / many-orbital spaces."""
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


def _rel(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)


@pytest.mark.parametrize("name", sorted(SMALL_SPACES))
def test_generic_kernels_on_the_small_spaces(torch_cuda, monkeypatch, name):
    import forte2_b200 as f2
    from oracle import checker
    monkeypatch.setenv("F2_GENERIC_KERNELS", "1")
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    c = synthetic.trial_vector(max(strings.ndet, 1))[: strings.ndet].copy()
    got = np.full_like(c, np.nan)
    b.Hamiltonian(c, got)
    want, _ = checker.sigma(space, E, H, V, c, dets=strings.make_determinants())
    assert _rel(got, want) <= TOL


LONG_LISTS = {
    # 3 alpha electrons in 45 orbitals: 129 single replacements per string, 946 parents per one-electron string
    "a3b1_45": (3, 1, 0, [[0] * 45], [], []),
    # 4 same-spin electrons in 20 orbitals (the example of ADVICE.md): 153 parents per two-electron string
    "a4b4_20": (4, 4, 0, [[0] * 20], [], []),
    # two GAS spaces with symmetry, 136 parents
    "gas_sym_19": (4, 2, 1, [[0, 1] * 5, [0, 1, 0, 1, 0, 1, 0, 1, 1]], [1, 0], [6, 5]),
}


@pytest.mark.parametrize("name", sorted(LONG_LISTS))
def test_spaces_with_long_lists(torch_cuda, name):
    import forte2_b200 as f2
    from oracle import sampled
    torch = torch_cuda
    space = LONG_LISTS[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    gen = torch.Generator(device="cuda").manual_seed(7)
    c = torch.randn(strings.ndet, dtype=torch.float64, device="cuda", generator=gen)
    c /= c.norm()
    sigma = torch.full_like(c, float("nan"))
    b.Hamiltonian(c, sigma)
    torch.cuda.synchronize()
    err, ns = sampled.sampled_rel_error(space, E, H, V, c, sigma, nsamp=800)
    print(name, strings.ndet, ns, err)
    assert ns > 100 and err <= TOL


# ---- ssb_kernel (same-spin term with the parent rows staged in shared memory, csrc/same_spin_staged.cuh): the default
# for spaces with several determinant blocks; forced everywhere with F2_SSB_ALWAYS, switched off with F2_NO_SSB
STAGED_SHAPES = {
    # 66 / 78 parents per (N-2)-electron string: beyond the staged kernel's 56 rows (ssg_kernel in every mode), one block
    "a3b2_13": (3, 2, 0, [[0] * 13], [], []),
    # the same with symmetry: per-group string lists, strings without parents in a group
    "a3b3_13_sym": (3, 3, 1, [[0, 1, 2, 3] * 3 + [0]], [], []),
    # 55 parents (CAS(18,18)'s shape, 56 rows, two CTAs of 8 warps), odd row lengths: the 8-byte copy path
    "a4b3_13": (4, 3, 0, [[0] * 13], [], []),
    # two GAS spaces, up to 91 / 105 parents: beyond the staged kernel's envelope (ssg_kernel in every mode)
    "gas_15": (3, 2, 0, [[0] * 7, [0] * 8], [0, 0], [5, 5]),
}


@pytest.mark.parametrize("mode", ["F2_SSB_ALWAYS", "F2_NO_SSB"])
@pytest.mark.parametrize("name", sorted(SMALL_SPACES) + sorted(STAGED_SHAPES))
def test_staged_same_spin_kernel(torch_cuda, monkeypatch, name, mode):
    import forte2_b200 as f2
    from oracle import checker, sampled
    torch = torch_cuda
    monkeypatch.setenv(mode, "1")
    space = SMALL_SPACES.get(name) or STAGED_SHAPES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    c = synthetic.trial_vector(max(strings.ndet, 1))[: strings.ndet].copy()
    got = np.full_like(c, np.nan)
    b.Hamiltonian(c, got)
    if strings.ndet <= 20000:
        want, _ = checker.sigma(space, E, H, V, c, dets=strings.make_determinants())
        assert _rel(got, want) <= TOL
    else:
        err, nsamp = sampled.sampled_rel_error(space, E, H, V, torch.from_numpy(c).cuda(), torch.from_numpy(got).cuda(), nsamp=1500)
        assert err <= TOL, (name, err)
