"""Parity at the sizes of BASELINE.json configs 3, 4 and 5, where no full oracle sigma can be computed: the kernel
instances that run there (t3g_kernel<4,1,0,384> at CAS(16,16), the long-row instance t3g_kernel<6,0,LONG,256> at
CAS(18,18), t3gp_kernel over the 225 blocks and 20-bit packed addresses of the 3-GAS / 20-orbital space, the
per-group K lists of ssg_kernel) are checked against the reference's definition of H through the sampled
Slater-rules checker oracle/sampled.py: for ~2000 seeded determinants I, sigma_I = sum_J <I|H|J> C_J rebuilt on the
host with C_J fetched from the GPU.  The checker itself is pinned against the reference's sigma in
tests/test_oracle.py.  One B200 holds every one of these spaces (C + sigma: 38 GB at CAS(18,18))."""
import numpy as np
import pytest

from forte2_b200 import synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _cas(n):
    return (n // 2, n // 2, 0, [[0] * n], [], [])


SPACES = {
    "cas14": _cas(14),                                                                # BASELINE configs[1]
    "cas16": _cas(16),                                                                # configs[2]
    "cas18": _cas(18),                                                                # configs[3]
    "gas20": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [8, 0, 0], [12, 16, 4]),        # configs[4]
    "gas20twin": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2]),
    "cas16_d2h": (8, 8, 0, [[i % 8 for i in range(16)]], [], []),
    "cas17_89": (8, 9, 0, [[0] * 17], [], []),                                        # na != nb, long rows, 384 threads
    "cas18_96": (9, 6, 0, [[0] * 18], [], []),                                        # 90 output rows, short rows of C
}


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


def _build(torch, name, memory_mb=65536):
    import forte2_b200 as f2
    space = SPACES[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    builder = f2.CISigmaBuilder(strings, E, H, V, 0)
    builder.set_memory(memory_mb)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(20261018)
    c = torch.randn(strings.ndet, dtype=torch.float64, device="cuda", generator=gen)
    c /= c.norm()
    return space, (E, H, V), strings, builder, c


@pytest.mark.parametrize("name,nsamp", [("cas14", 2000), ("cas16", 2000), ("cas16_d2h", 2000), ("gas20twin", 2000),
                                        ("cas17_89", 1000), ("cas18_96", 1000), ("cas18", 1500), ("gas20", 1500)])
def test_sigma_at_sampled_determinants(torch_cuda, name, nsamp):
    torch = torch_cuda
    from oracle import sampled
    space, ints, strings, builder, c = _build(torch, name)
    assert strings.ndet == sampled.layout(space).ndet
    sigma = torch.full_like(c, float("nan"))
    builder.Hamiltonian(c, sigma)
    torch.cuda.synchronize()
    err, n = sampled.sampled_rel_error(space, *ints, c, sigma, nsamp=nsamp)
    print(f"{name}: ndet={strings.ndet} blocks={strings.nblocks} form={builder.t3_shape()['form']} "
          f"sampled={n} rel L2 error={err:.2e}")
    assert n >= nsamp // 2 and err <= TOL
    del sigma, c, builder
    torch.cuda.empty_cache()


@pytest.mark.parametrize("env", [{"F2_T3G_WIDE": "1"}, {"F2_T3G_NSM": "1000"}, {"F2_T3G_NSM": "1000", "F2_T3G_WIDE": "1"}])
def test_six_row_tile_instances_of_the_gather_kernel(torch_cuda, monkeypatch, env):
    """90 output rows (9 alpha electrons in 18 orbitals) select t3g_kernel<6, ...>: 8-column items on 12 warps by default,
    16-column items on 8 warps with F2_T3G_WIDE, each with the whole row of C staged or (F2_T3G_NSM) partly gathered from L2"""
    torch = torch_cuda
    from oracle import sampled
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    space, ints, strings, builder, c = _build(torch, "cas18_96")
    sigma = torch.full_like(c, float("nan"))
    builder.Hamiltonian(c, sigma)
    torch.cuda.synchronize()
    err, n = sampled.sampled_rel_error(space, *ints, c, sigma, nsamp=600)
    assert builder.t3_shape()["form"] == "gather" and err <= TOL
    del sigma, c, builder
    torch.cuda.empty_cache()


def test_sampled_check_sees_a_wrong_sigma(torch_cuda):
    """the checker is not vacuous: one wrong integral, or sigma of a neighbouring vector, fails it by orders of magnitude"""
    torch = torch_cuda
    from oracle import sampled
    space, (E, H, V), strings, builder, c = _build(torch, "cas14")
    sigma = torch.empty_like(c)
    builder.Hamiltonian(c, sigma)
    V2 = V.copy()
    V2[1, 2, 3, 4] += 1e-6
    err, _ = sampled.sampled_rel_error(space, E, H, V2, c, sigma, nsamp=400)
    assert err > 1e-9
    err, _ = sampled.sampled_rel_error(space, E, H, V, c, sigma * (1 + 1e-10), nsamp=400)
    assert 5e-11 < err < 2e-10


@pytest.mark.parametrize("name,world,nsamp", [("cas16", 8, 1000), ("cas18", 8, 1000), ("gas20twin", 2, 1000),
                                              ("gas20", 8, 1000)])
def test_rows_form_at_size_emulated(torch_cuda, name, world, nsamp):
    """the sharded (rows-form) kernel instances at their natural sizes: every rank's f2_sigma_apply_rows run in turn on
    this GPU (tests/rows_emulation.py), the assembled sigma checked at sampled determinants"""
    torch = torch_cuda
    from oracle import sampled
    from rows_emulation import rows_form_sigma
    space, ints, strings, builder, c = _build(torch, name, memory_mb=16384)
    sigma = rows_form_sigma(torch, strings, builder, c, world, nan_fill=False)
    torch.cuda.synchronize()
    err, n = sampled.sampled_rel_error(space, *ints, c, sigma, nsamp=nsamp)
    print(f"{name} rows form, world {world}: sampled={n} rel L2 error={err:.2e}")
    assert err <= TOL
    del sigma, c, builder
    torch.cuda.empty_cache()
