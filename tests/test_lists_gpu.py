"""Replacement lists built by CUDA kernels (forte2_b200/csrc/lists_device.cu) against the host builder (strings.cc, itself
bit-exact against the reference's lists in tests/test_strings_cpu.py): every array the sigma and RDM kernels read, both
spins, identical element by element."""
import numpy as np
import pytest

from cases import SMALL_SPACES
import forte2_b200 as f2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


KINDS = ["dev_as_ptr", "dev_as_ent", "dev_bs_ent", "dev_h2_ptr", "dev_h2_ent", "dev_cls", "dev_loc"]
SPACES = dict(SMALL_SPACES)
SPACES["cas1010"] = (5, 5, 0, [[0] * 10], [], [])
SPACES["gas_three_12"] = (4, 3, 0, [[0] * 4, [0] * 5, [0] * 3], [2, 0, 0], [6, 7, 3])
SPACES["sym8_10"] = (4, 4, 3, [[0, 1, 2, 3, 4, 5, 6, 7, 0, 3]], [], [])


@pytest.mark.parametrize("name", sorted(SPACES))
def test_device_built_lists_equal_host_built_lists(torch_cuda, name):
    space = SPACES[name]
    host, dev = f2.CIStrings(*space), f2.CIStrings(*space)
    host.build_lists(None)
    dev.build_lists(0)
    for kind in KINDS:
        for spin in (0, 1):
            a, b = host.export(kind, spin), dev.export(kind, spin)
            assert a.shape == b.shape and np.array_equal(a, b), (kind, spin)


def test_sigma_builder_builds_its_lists_on_the_device(torch_cuda, monkeypatch):
    """a builder made from fresh strings runs the device list builder; F2_HOST_LISTS selects the host builder; same sigma"""
    torch = torch_cuda
    from forte2_b200 import synthetic
    space = SPACES["gas343_54"]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    out = []
    for host_lists in (False, True):
        if host_lists:
            monkeypatch.setenv("F2_HOST_LISTS", "1")
        strings = f2.CIStrings(*space)
        b = f2.CISigmaBuilder(strings, E, H, V, 0)
        c = torch.from_numpy(synthetic.trial_vector(strings.ndet)).cuda()
        s = torch.empty_like(c)
        b.Hamiltonian(c, s)
        out.append(s.cpu().numpy())
    assert np.linalg.norm(out[0] - out[1]) <= 1e-13 * np.linalg.norm(out[1])
