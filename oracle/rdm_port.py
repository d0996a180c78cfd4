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
