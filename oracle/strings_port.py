"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT (only tests/, smoke() and bench.py's cpu_baseline leg
import this).

Pure-Python restatement, for SMALL spaces, of how the reference lays out the CI vector:
  occupation patterns   forte2/ci/ci_occupation.cc:35-142   (recursive ORMAS generation)
  string order          forte2/ci/ci_strings_makers.cc:13-73 (next_permutation per GAS space, Cartesian
                        product with the last space fastest, stable bucket by irrep)
  classes and blocks    forte2/ci/ci_string_class.cc:39-65
  determinant order     forte2/ci/ci_strings.cc:152-160,186-203
Parity pinned against oracle/_ref and tests/golden/lists_*.npz in tests/test_oracle.py.
"""
from __future__ import annotations

from itertools import combinations, product

import numpy as np


def occupation_patterns(na, nb, gas_min, gas_max, gas_size):
    lo, hi, size = [0] * 6, [0] * 6, [0] * 6
    nspaces = 0
    for n in range(6):
        sz = gas_size[n] if n < len(gas_size) else 0
        if not sz:
            continue
        emax = min(2 * sz, na + nb)
        if len(gas_max) > n:
            emax = min(gas_max[n], emax)
        hi[n], lo[n], size[n] = emax, (gas_min[n] if len(gas_min) > n else 0), sz
        nspaces += 1
    if nspaces == 0:
        return 1, [(0,) * 6], [(0,) * 6], [(0, 0)]
    alpha, beta, pairs = [], [], []

    def intern(lst, o):
        if o not in lst:
            lst.append(o)
        return lst.index(o)

    def walk(space, ra, rb, oa, ob):
        if space == nspaces - 1:
            if lo[space] <= ra + rb <= hi[space] and max(ra, rb) <= size[space] and min(ra, rb) >= 0:
                oa, ob = list(oa), list(ob)
                oa[space], ob[space] = ra, rb
                pairs.append((intern(alpha, tuple(oa)), intern(beta, tuple(ob))))
            return
        for xa in range(max(lo[space] - size[space], 0), min(hi[space], size[space]) + 1):
            for xb in range(max(lo[space] - xa, 0), min(hi[space] - xa, size[space]) + 1):
                if lo[space] <= xa + xb <= hi[space]:
                    oa2, ob2 = list(oa), list(ob)
                    oa2[space], ob2[space] = xa, xb
                    walk(space + 1, ra - xa, rb - xb, oa2, ob2)

    walk(0, na, nb, [0] * 6, [0] * 6)
    return nspaces, alpha, beta, pairs


def _space_strings(start, m, k):
    """k electrons in orbitals start..start+m-1, lexicographic in (b0, b1, ...) with empty < occupied."""
    if k < 0 or k > m:
        return None
    out = []
    for occ in combinations(range(m), k):
        out.append(tuple(1 if i in occ else 0 for i in range(m)))
    out.sort()
    return [sum(1 << (start + i) for i in range(m) if bits[i]) for bits in out]


def strings_by_class(occs, gas_size, mo_sym, nirrep, ngas):
    starts = np.concatenate([[0], np.cumsum(gas_size)]).astype(int)
    classes = []
    for occ in occs:
        per_space = [_space_strings(int(starts[g]), gas_size[g], occ[g]) for g in range(ngas)]
        buckets = [[] for _ in range(nirrep)]
        if all(p is not None for p in per_space):
            for combo in product(*per_space):  # last space fastest
                w = 0
                for x in combo:
                    w |= x
                h = 0
                for p in range(len(mo_sym)):
                    if (w >> p) & 1:
                        h ^= mo_sym[p]
                buckets[h].append(w)
        classes.extend(buckets)
    return classes


def determinants(na, nb, symmetry, orbital_symmetry, gas_min, gas_max):
    """(ndet, 2) uint64 array of (alpha word, beta word) in the reference's CI-vector order."""
    gas_size = [len(s) for s in orbital_symmetry]
    mo_sym = [x for s in orbital_symmetry for x in s]
    nirrep = max([1] + [abs(s + 1) for s in mo_sym])
    ngas, aocc, bocc, pairs = occupation_patterns(na, nb, gas_min, gas_max, gas_size)
    gas_size = gas_size + [0] * (ngas - len(gas_size))
    A = strings_by_class(aocc, gas_size, mo_sym, nirrep, ngas)
    B = strings_by_class(bocc, gas_size, mo_sym, nirrep, ngas)
    dets = []
    for ia, ib in pairs:
        for h in range(nirrep):
            for a in A[ia * nirrep + h]:
                for b in B[ib * nirrep + (h ^ symmetry)]:
                    dets.append((a, b))
    return np.array(dets, dtype=np.uint64).reshape(-1, 2)


class Layout:
    """String tables and determinant blocks of a space WITHOUT enumerating its determinants (the sampled checker,
    oracle/sampled.py, runs at 10^9 determinants): per spin the strings class by class in address order, and the
    block table (class_A, class_B, offset, nA, nB) in determinant_classes order (ci_string_class.cc:57-65,
    ci_strings.cc:152-160)."""

    def __init__(self, na, nb, symmetry, orbital_symmetry, gas_min, gas_max):
        gas_size = [len(s) for s in orbital_symmetry]
        mo_sym = [x for s in orbital_symmetry for x in s]
        self.norb, self.na, self.nb = len(mo_sym), na, nb
        nirrep = max([1] + [abs(s + 1) for s in mo_sym])
        ngas, aocc, bocc, pairs = occupation_patterns(na, nb, gas_min, gas_max, gas_size)
        gas_size = gas_size + [0] * (ngas - len(gas_size))
        self.classes = [strings_by_class(occ, gas_size, mo_sym, nirrep, ngas) for occ in (aocc, bocc)]
        self.sorted_words, self.word_class, self.word_addr = [], [], []
        for cls in self.classes:
            words = np.array([w for c in cls for w in c], dtype=np.uint64)
            cid = np.array([k for k, c in enumerate(cls) for _ in c], dtype=np.int64)
            adr = np.array([i for c in cls for i in range(len(c))], dtype=np.int64)
            order = np.argsort(words, kind="stable")
            self.sorted_words.append(words[order])
            self.word_class.append(cid[order])
            self.word_addr.append(adr[order])
        ncA, ncB = len(self.classes[0]), len(self.classes[1])
        self.block_off = np.full((ncA, ncB), -1, dtype=np.int64)
        self.blocks = []
        off = 0
        for ia, ib in pairs:
            for h in range(nirrep):
                cA, cB = ia * nirrep + h, ib * nirrep + (h ^ symmetry)
                nA, nB = len(self.classes[0][cA]), len(self.classes[1][cB])
                self.blocks.append((cA, cB, off, nA, nB))
                if nA * nB > 0:
                    self.block_off[cA, cB] = off
                off += nA * nB
        self.ndet = off
        self.ncols = np.array([len(c) for c in self.classes[1]], dtype=np.int64)

    def lookup(self, spin, words):
        """(class, address) of every occupation word, class -1 where the string is not in the space"""
        words = np.asarray(words, dtype=np.uint64)
        table = self.sorted_words[spin]
        if table.size == 0:
            return np.full(words.shape, -1, dtype=np.int64), np.zeros(words.shape, dtype=np.int64)
        pos = np.minimum(np.searchsorted(table, words), table.size - 1)
        hit = table[pos] == words
        return np.where(hit, self.word_class[spin][pos], -1), np.where(hit, self.word_addr[spin][pos], 0)

    def address(self, a_words, b_words):
        """CI-vector address of (alpha word, beta word), -1 if the determinant is outside the space"""
        cA, iA = self.lookup(0, a_words)
        cB, iB = self.lookup(1, b_words)
        ok = (cA >= 0) & (cB >= 0)
        off = np.where(ok, self.block_off[np.maximum(cA, 0), np.maximum(cB, 0)], -1)
        return np.where(off >= 0, off + iA * self.ncols[np.maximum(cB, 0)] + iB, -1)

    def determinant(self, addresses):
        """(alpha word, beta word) at the given CI-vector addresses"""
        addresses = np.asarray(addresses, dtype=np.int64)
        starts = np.array([b[2] for b in self.blocks], dtype=np.int64)
        sizes = np.array([b[3] * b[4] for b in self.blocks], dtype=np.int64)
        live = np.nonzero(sizes > 0)[0]
        bi = live[np.searchsorted(starts[live], addresses, side="right") - 1]
        a_out, b_out = np.zeros(addresses.shape, dtype=np.uint64), np.zeros(addresses.shape, dtype=np.uint64)
        for k, (i, adr) in enumerate(zip(bi, addresses)):
            cA, cB, off, nA, nB = self.blocks[i]
            r, c = divmod(int(adr - off), nB)
            a_out[k], b_out[k] = self.classes[0][cA][r], self.classes[1][cB][c]
        return a_out, b_out
