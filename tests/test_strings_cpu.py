"""CPU tests of the host string model through the C ABI: bit-exact against the committed golden
fixtures (tests/golden, generated from the reference by make_golden.py), against the occupation
tables hard-coded in the reference's own tests, and against the live reference library when it has
been built (oracle/_ref)."""
import os

import numpy as np
import pytest

from cases import SMALL_SPACES
import forte2_b200 as f2
from forte2_b200.ci_strings import KINDS

GOLD = os.path.join(os.path.dirname(__file__), "golden")
LIST_FIXTURES = sorted(f[len("lists_"):-4] for f in os.listdir(GOLD) if f.startswith("lists_"))


@pytest.mark.parametrize("name", LIST_FIXTURES)
def test_lists_bit_exact_vs_golden(name):
    g = np.load(os.path.join(GOLD, f"lists_{name}.npz"))
    s = f2.CIStrings(*SMALL_SPACES[name])
    info = [s.na, s.nb, s.nas, s.nbs, s.ndet, s.norb, s.nirrep, s.ngas_spaces, s.nblocks, s.symmetry]
    assert info == [int(x) for x in g["info"]]
    assert np.array_equal(s.blocks(), g["blocks"])
    assert np.array_equal(s.occupations(0), g["occ_a"])
    assert np.array_equal(s.occupations(1), g["occ_b"])
    assert np.array_equal(s.occupations(2), g["occ_pairs"])
    for kind in KINDS:
        for spin in (0, 1):
            assert np.array_equal(s.export(kind, spin).reshape(-1), g[f"{kind}_{spin}"]), (kind, spin)


@pytest.mark.parametrize("name", sorted(SMALL_SPACES))
def test_lists_bit_exact_vs_live_reference(ref, name):
    space = SMALL_SPACES[name]
    r = ref.RefCIStrings(*space)
    s = f2.CIStrings(*space)
    assert (s.nas, s.nbs, s.ndet, s.nblocks, s.nirrep) == (r.nas, r.nbs, r.ndet, r.nblocks, r.nirrep)
    assert np.array_equal(s.blocks(), r.blocks())
    for kind in KINDS:
        for spin in (0, 1):
            assert np.array_equal(s.export(kind, spin).reshape(-1), r.export(kind, spin).reshape(-1)), (kind, spin)
    assert np.array_equal(s.make_determinants(), r.determinants())


def test_reference_known_answers_singlet():
    # /root/reference/tests/ci/test_ci_strings.py:22-50
    s = f2.CIStrings(4, 4, 0, [[0], [0, 0], [0, 0, 0]], [1, 4, 2], [2, 4, 3])
    assert (s.na, s.nb, s.ngas_spaces, s.gas_size) == (4, 4, 3, [1, 2, 3])
    assert s.gas_alpha_occupations == [[0, 2, 2, 0, 0, 0], [1, 2, 1, 0, 0, 0]]
    assert s.gas_beta_occupations == [[1, 2, 1, 0, 0, 0], [0, 2, 2, 0, 0, 0]]
    assert s.gas_occupations == [(0, 0), (1, 1), (1, 0)]


def test_reference_known_answers_triplet():
    # /root/reference/tests/ci/test_ci_strings.py:53-101
    s = f2.CIStrings(5, 3, 0, [[0], [0, 0], [0, 0, 0]], [0, 3], [])
    assert s.gas_alpha_occupations == [[0, 2, 3, 0, 0, 0], [1, 1, 3, 0, 0, 0], [1, 2, 2, 0, 0, 0]]
    assert s.gas_beta_occupations == [[0, 1, 2, 0, 0, 0], [0, 2, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0],
                                      [1, 2, 0, 0, 0, 0]]
    assert s.gas_occupations == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 1), (2, 0), (2, 1), (1, 3), (2, 2),
                                 (2, 3)]


def test_reference_symmetry_uses_all_set_bits():
    # /root/reference/tests/ci/test_ci_strings.py:104-131 (64 orbitals, 6 GAS spaces)
    orbsym = [[i % 4 for i in range(59)], [1], [2], [3], [1], [2]]
    s = f2.CIStrings(3, 3, 3, orbsym, [1] * 6, [1] * 6)
    mo = [x for g in orbsym for x in g]
    dets = s.make_determinants()
    assert len(dets) > 0
    for a, b in dets:
        a, b = int(a), int(b)
        assert ((a >> 63) & 1) != ((b >> 63) & 1)
        sym = 0
        for p in range(64):
            if (a >> p) & 1:
                sym ^= mo[p]
            if (b >> p) & 1:
                sym ^= mo[p]
        assert sym == 3


def test_input_validation():
    # /root/reference/tests/ci/test_ci_strings.py:6-19
    with pytest.raises(ValueError):
        f2.CIStrings(-1, 0, 0, [[]], [], [])
    with pytest.raises(ValueError):
        f2.CIStrings(-1, -1, 0, [[]], [], [])
    with pytest.raises(ValueError):
        f2.CIStrings(1, -1, 0, [[]], [], [])
    with pytest.raises(TypeError):
        f2.CIStrings(1, 1, 0, [0], [], [])
    with pytest.raises(ValueError):
        f2.CIStrings(1, 1, 0, [[0, 0]], [1, 1], [])  # more gas_min entries than spaces


def test_appendix_b_string_order():
    # address 0 has the highest orbitals occupied; bit p = orbital p
    s = f2.CIStrings(2, 1, 0, [[0] * 4], [], [])
    words = s.export("strings", 0)
    assert list(words[2:]) == [0b1100, 0b1010, 0b0110, 0b1001, 0b0101, 0b0011]
    assert s.export("1h2", 0).tolist()[:3] == [[0, 0, 0, 1, 2, 0], [0, 0, 0, 1, 1, 1], [0, 0, 0, 1, 0, 3]]
    assert s.export("2h", 0).tolist() == [[0, 0, 0, 1, 3, 2, 0], [0, 0, 0, 1, 3, 1, 1], [0, 0, 0, 1, 2, 1, 2],
                                           [0, 0, 0, 1, 3, 0, 3], [0, 0, 0, 1, 2, 0, 4], [0, 0, 0, 1, 1, 0, 5]]


def test_determinant_addressing_round_trip():
    s = f2.CIStrings(4, 3, 2, [[0, 1, 2], [3, 0, 1, 2]], [2, 0], [6, 7])
    dets = s.make_determinants()
    assert dets.shape == (s.ndet, 2)
    for addr in range(0, s.ndet, 7):
        d = s.determinant(addr)
        assert d == (int(dets[addr, 0]), int(dets[addr, 1]))
        assert s.determinant_index(d) == addr
    with pytest.raises(IndexError):
        s.determinant_index((0, 0))


def test_large_space_sizes():
    # the BASELINE configs' sizes, from the string model alone
    s = f2.CIStrings(7, 7, 0, [[0] * 14], [], [])
    assert (s.nas, s.ndet) == (3432, 11778624)
    s = f2.CIStrings(10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])
    assert (s.nas, s.nbs, s.ndet, s.nblocks) == (29932, 29932, 22065484, 36)


def test_shard_layout_index_maps():
    """forte2_b200/dist.py Layout: local <-> CI-vector address maps of the alpha-row partition are
    inverse to each other and the ranks' pieces tile the vector exactly once."""
    import forte2_b200 as f2
    from forte2_b200.dist import Layout
    for name in ["cas66", "gas343_54", "cas66_d2h", "cas75_odd", "gas2_sym4"]:
        strings = f2.CIStrings(*SMALL_SPACES[name])
        for world in (1, 2, 3, 5):
            seen = np.zeros(strings.ndet, dtype=np.int64)
            for rank in range(world):
                lay = Layout(strings.blocks(), rank, world)
                loc = np.arange(lay.local_size)
                glob = lay.global_of_local(loc)
                seen[glob] += 1
                assert np.array_equal(lay.local_of_global(glob), loc)
                others = np.setdiff1d(np.arange(strings.ndet), glob)
                assert np.all(lay.local_of_global(others) == -1)
            assert np.all(seen == 1)


@pytest.mark.parametrize("name,world", [("gas343_54", 8), ("cas66_d2h", 5), ("gas2_sym4", 8), ("cas66", 7),
                                        ("one_beta_three_alpha", 4)])
def test_rows_form_exchange_plans_are_consistent(name, world):
    """forte2_b200.dist.ShardedSigma._cols_plan / _exchange_plan for many ranks and classes smaller than the
    world (ranks that own no row or no column of a block): what rank r packs for q is exactly what q expects from
    r, and replaying the all-to-all in one process assembles the column slabs of a known vector and returns the
    alpha same-spin partials to the row owners."""
    import torch
    from forte2_b200.dist import ShardedSigma

    from forte2_b200.dist import Layout

    strings = f2.CIStrings(*SMALL_SPACES[name])
    blocks = strings.blocks()
    slot = max(Layout(blocks, q, world).local_size for q in range(world))

    class NoEngine:  # the plans need the layout only
        def rows_form(self, rank, world):
            return True, slot

    full = torch.arange(strings.ndet, dtype=torch.float64) + 1.0
    ranks = [ShardedSigma(rank=r, world=world, engine=NoEngine(), blocks=blocks, device=torch.device("cpu"),
                          form="rows") for r in range(world)]
    assert sum(sh.local_size for sh in ranks) == strings.ndet
    local = [sh.scatter_full(full) for sh in ranks]
    # ---- step 1 of the rows form: column slabs
    for q, shq in enumerate(ranks):
        shq._xcols.fill_(float("nan"))
    for r, shr in enumerate(ranks):
        for q, (sends, _, _, _) in enumerate(shr._cols_plan()):
            pieces = [local[r][loc:loc + n * nB].view(n, nB)[:, cl:cl + nc].reshape(-1) for loc, n, nB, cl, nc in sends]
            packed = torch.cat(pieces) if pieces else torch.zeros(0, dtype=torch.float64)
            recvs = ranks[q]._cols_plan()[r][1]
            assert packed.numel() == sum(m for _, m in recvs), (r, q)
            pos = 0
            for start, m in recvs:
                ranks[q]._xcols[start:start + m] = packed[pos:pos + m]
                pos += m
    for q, shq in enumerate(ranks):
        for (off, nA, nB), (cl, nc, yoff) in zip(shq.layout.blocks, shq.layout.cols):
            if nA * nB > 0:
                want = full[off:off + nA * nB].view(nA, nB)[:, cl:cl + nc].reshape(-1)
                assert torch.equal(shq._xcols[yoff:yoff + nA * nc], want), (name, q)
    # ---- step 3: the partials [all rows][my columns] go back to the row owners and are added
    got = [torch.zeros(sh.local_size, dtype=torch.float64) for sh in ranks]
    for r, shr in enumerate(ranks):
        shr._yexch = shr._xcols.clone()  # pretend the alpha same-spin operator is the identity
        for q, (send_slices, _, _, _) in enumerate(shr._exchange_plan()):
            packed = torch.cat([shr._yexch[lo:hi] for lo, hi in send_slices]) if send_slices else torch.zeros(0)
            adds = ranks[q]._exchange_plan()[r][3]
            assert packed.numel() == sum(n * nc for _, n, _, _, nc, _ in adds), (r, q)
            for loc, n, nB, cl, nc, pos in adds:
                got[q][loc:loc + n * nB].view(n, nB)[:, cl:cl + nc] += packed[pos:pos + n * nc].view(n, nc)
    for q in range(world):
        assert torch.equal(got[q], local[q]), (name, q)
