"""TEST INFRASTRUCTURE (never imported by the product): reduced density matrices from their definitions, a plain
Python/numpy restatement over the determinant list, for small spaces only.  The checker of the next widening step
(1- and 2-RDMs on the GPU); pinned against the reference's own
CISigmaBuilder::compute_*rdm (forte2/ci/ci_sigma_builder_1rdm.cc:9-84, ci_sigma_builder_2rdm.cc:45-282) in
tests/test_oracle.py.

Conventions (those of the reference):
  a_1rdm[p, q]        = <L| a+_p a_q |R>                      (alpha; b_1rdm the same for beta)
  aa_2rdm[pq, rs]     = <L| a+_p a+_q a_s a_r |R>              p > q, r > s, pair index p (p - 1) / 2 + q
  ab_2rdm[u, v, x, y] = <L| a+_u(alpha) a+_v(beta) a_y(beta) a_x(alpha) |R>, no cross-spin phase
  sf_1rdm, sf_2rdm, sf_2cumulant as assembled in ci_sigma_builder_1rdm.cc:71-81, ci_sigma_builder_2rdm.cc:203-279
`dets` is the (ndet, 2) uint64 array of (alpha, beta) occupation words in CI-vector order (bit p = orbital p);
the sign of removing or adding orbital n is (-1)^(number of set bits below n) (determinant/bitarray.hpp:575-583)."""
import numpy as np


def _sign(word, n):
    return -1.0 if bin(word & ((1 << n) - 1)).count("1") & 1 else 1.0


def _occ(word, norb):
    return [p for p in range(norb) if (word >> p) & 1]


def _index(dets):
    return {(int(a), int(b)): i for i, (a, b) in enumerate(np.asarray(dets, dtype=np.uint64))}


def s_1rdm(dets, norb, left, right, spin):
    """spin 0 = alpha, 1 = beta"""
    where = _index(dets)
    out = np.zeros((norb, norb))
    for (a, b), i in where.items():
        w = a if spin == 0 else b
        for q in _occ(w, norb):
            sq = _sign(w, q)
            k = w ^ (1 << q)
            for p in range(norb):
                if (k >> p) & 1:
                    continue
                j = where.get((k | (1 << p), b) if spin == 0 else (a, k | (1 << p)))
                if j is not None:
                    out[p, q] += sq * _sign(k, p) * left[j] * right[i]
    return out


def pair_index_gt(p, q):
    return p * (p - 1) // 2 + q


def ss_2rdm(dets, norb, left, right, spin):
    where = _index(dets)
    npairs = norb * (norb - 1) // 2
    out = np.zeros((npairs, npairs))
    for (a, b), i in where.items():
        w = a if spin == 0 else b
        occ = _occ(w, norb)
        for r in occ:
            for s in occ:
                if s >= r:
                    continue
                s_right = _sign(w, s) * _sign(w ^ (1 << s), r)  # K = a_r a_s I, r > s
                k = w ^ (1 << r) ^ (1 << s)
                for p in range(norb):
                    if (k >> p) & 1:
                        continue
                    for q in range(p):
                        if (k >> q) & 1:
                            continue
                        par = k | (1 << p) | (1 << q)
                        j = where.get((par, b) if spin == 0 else (a, par))
                        if j is None:
                            continue
                        s_left = _sign(par, q) * _sign(par ^ (1 << q), p)
                        out[pair_index_gt(p, q), pair_index_gt(r, s)] += s_left * s_right * left[j] * right[i]
    return out


def ab_2rdm(dets, norb, left, right):
    where = _index(dets)
    out = np.zeros((norb,) * 4)
    for (a, b), i in where.items():
        for x in _occ(a, norb):
            ka, sx = a ^ (1 << x), _sign(a, x)
            for y in _occ(b, norb):
                kb, sy = b ^ (1 << y), _sign(b, y)
                for u in range(norb):
                    if (ka >> u) & 1:
                        continue
                    su = _sign(ka, u)
                    for v in range(norb):
                        if (kb >> v) & 1:
                            continue
                        j = where.get((ka | (1 << u), kb | (1 << v)))
                        if j is not None:
                            out[u, v, x, y] += su * _sign(kb, v) * sx * sy * left[j] * right[i]
    return out


def sf_1rdm(dets, norb, left, right):
    return s_1rdm(dets, norb, left, right, 0) + s_1rdm(dets, norb, left, right, 1)


def sf_2rdm(dets, norb, left, right):
    ab = ab_2rdm(dets, norb, left, right)
    out = ab + ab.transpose(1, 0, 3, 2)
    for spin in (0, 1):
        ss = ss_2rdm(dets, norb, left, right, spin)
        for p in range(1, norb):
            for q in range(p):
                for r in range(1, norb):
                    for s in range(r):
                        e = ss[pair_index_gt(p, q), pair_index_gt(r, s)]
                        out[p, q, r, s] += e
                        out[q, p, r, s] -= e
                        out[p, q, s, r] -= e
                        out[q, p, s, r] += e
    return out


def sf_2cumulant(dets, norb, left, right):
    g1 = sf_1rdm(dets, norb, left, right)
    return (sf_2rdm(dets, norb, left, right) - np.einsum("pr,qs->pqrs", g1, g1)
            + 0.5 * np.einsum("ps,qr->pqrs", g1, g1))


# ---- any spin block up to three-body, from the definition (the 3-RDM blocks of ci_sigma_builder_3rdm.cc:86-318) -------------
def _op_index(orbs):
    """index of a strictly decreasing tuple of orbitals: p; p (p-1)/2 + q; p (p-1) (p-2)/6 + q (q-1)/2 + r
    (helpers/indexing.hpp: pair_index_gt, triplet_index_gt)"""
    if len(orbs) == 0:
        return 0
    if len(orbs) == 1:
        return orbs[0]
    if len(orbs) == 2:
        p, q = orbs
        return p * (p - 1) // 2 + q
    p, q, r = orbs
    return p * (p - 1) * (p - 2) // 6 + q * (q - 1) // 2 + r


def _nops(norb, level):
    from math import comb
    return comb(norb, level)


def _removals(word, norb, level):
    """every way to annihilate `level` electrons of `word` (smallest orbital first): (remaining word, op index of the
    decreasing orbital tuple, sign)"""
    from itertools import combinations
    out = []
    for orbs in combinations(_occ(word, norb), level):  # increasing
        w, sign = word, 1.0
        for o in orbs:
            sign *= _sign(w, o)
            w ^= 1 << o
        out.append((w, _op_index(tuple(reversed(orbs))), sign))
    return out


def general_rdm(dets, norb, left, right, la, lb):
    """<L| (la alpha creators) (lb beta creators) (lb beta annihilators) (la alpha annihilators) |R> as the matrix
    [op_a * P_b + op_b, op_a' * P_b + op_b'] over decreasing orbital tuples: (1,0) a_1rdm, (0,1) b_1rdm, (2,0) aa_2rdm,
    (1,1) ab_2rdm, (3,0) aaa_3rdm, (2,1) aab_3rdm [(uv), w, (xy), z], (1,2) abb_3rdm [u, (vw), x, (yz)], (0,3) bbb_3rdm"""
    Pa, Pb = _nops(norb, la), _nops(norb, lb)
    out = np.zeros((Pa * Pb, Pa * Pb))
    groups = {}
    for vec, side in ((left, 0), (right, 1)):
        for (a, b), c in zip(np.asarray(dets, dtype=np.uint64), vec):
            if c == 0.0:
                continue
            for ka, opa, sa in _removals(int(a), norb, la):
                for kb, opb, sb in _removals(int(b), norb, lb):
                    groups.setdefault((ka, kb), ([], []))[side].append((opa * Pb + opb, sa * sb * c))
    for lhs, rhs in groups.values():
        for i, x in lhs:
            for j, y in rhs:
                out[i, j] += x * y
    return out


def sf_3rdm(dets, norb, left, right):
    """spin-free 3-RDM G3[p,q,r,s,t,u] = sum over spins of <L| a+_p a+_q a+_r a_u a_t a_s |R> (compute_sf_3rdm,
    ci_sigma_builder_3rdm.cc:320-495), straight from the definition: every spin assignment (x, y, z) of the three
    creator/annihilator pairs, each as a generalised spin-block element with its reordering sign"""
    from itertools import permutations, product
    n = norb
    out = np.zeros((n,) * 6)
    blocks = {(la, 3 - la): general_rdm(dets, n, left, right, la, 3 - la) for la in range(4)}

    def perm_sign(seq):
        seq, sign = list(seq), 1.0
        for i in range(len(seq)):
            for j in range(len(seq) - 1 - i):
                if seq[j] < seq[j + 1]:
                    seq[j], seq[j + 1] = seq[j + 1], seq[j]
                    sign = -sign
        return sign, tuple(seq)

    def element(cre_a, cre_b, ann_a, ann_b):
        # <L| a+(cre_a...) b+(cre_b...) b(ann_b reversed) a(ann_a reversed) |R>, each list in the given order
        if len(set(cre_a)) < len(cre_a) or len(set(cre_b)) < len(cre_b) or len(set(ann_a)) < len(ann_a) \
                or len(set(ann_b)) < len(ann_b):
            return 0.0
        s1, ca = perm_sign(cre_a)
        s2, cb = perm_sign(cre_b)
        s3, aa = perm_sign(ann_a)
        s4, ab = perm_sign(ann_b)
        la, lb = len(ca), len(cb)
        Pb = _nops(n, lb)
        return s1 * s2 * s3 * s4 * blocks[(la, lb)][_op_index(ca) * Pb + _op_index(cb), _op_index(aa) * Pb + _op_index(ab)]

    for spins in product((0, 1), repeat=3):
        for p, q, r, s, t, u in product(range(n), repeat=6):
            cre, ann = (p, q, r), (s, t, u)
            # operator a+_p(x) a+_q(y) a+_r(z) a_u(z) a_t(y) a_s(x): move the alpha operators to the outside
            cre_a = [o for o, sp in zip(cre, spins) if sp == 0]
            cre_b = [o for o, sp in zip(cre, spins) if sp == 1]
            ann_a = [o for o, sp in zip(ann, spins) if sp == 0]
            ann_b = [o for o, sp in zip(ann, spins) if sp == 1]
            # sign of sorting the creators (alpha first) and, identically, the annihilators: they cancel
            out[p, q, r, s, t, u] += element(cre_a, cre_b, ann_a, ann_b)
    return out


def trdm1(dets_left, dets_right, norb, left, right, spin):
    """transition 1-RDM <L| a+_p a_q |R> of spin 0 / 1 between two CI spaces with the same orbitals and electron numbers
    (compute_s_1trdm, forte2/ci/ci_sigma_builder_1trdm.cc:11-113), from the definition"""
    where = _index(dets_left)
    out = np.zeros((norb, norb))
    for (a, b), c in zip(np.asarray(dets_right, dtype=np.uint64), right):
        a, b = int(a), int(b)
        w = a if spin == 0 else b
        for q in _occ(w, norb):
            sq = _sign(w, q)
            k = w ^ (1 << q)
            for p in range(norb):
                if (k >> p) & 1:
                    continue
                j = where.get((k | (1 << p), b) if spin == 0 else (a, k | (1 << p)))
                if j is not None:
                    out[p, q] += sq * _sign(k, p) * left[j] * c
    return out
