"""Synthetic active-space integrals used by the parity tests and bench.py.

The hand-off format is the reference's: scalar E, H[n,n] and V[n,n,n,n] in physicist notation
<pq|rs> (forte2/jkbuilder/mointegrals.py:40,84), C-contiguous float64.  Both variants are 8-fold
symmetric (real orbitals), which is what forte2's density-fitted integrals always are and what the
reference's Knowles-Handy algorithm requires (forte2/ci/ci_sigma_builder.cc:172-184).
"""
from __future__ import annotations

import numpy as np

SEED = 20261017


def _pair_expand(g: np.ndarray, n: int) -> np.ndarray:
    """V[p,q,r,s] = g[pair(p,r), pair(q,s)] with pair(i,j) = max(max+1)/2 + min  (chemist (pr|qs))."""
    idx = np.arange(n)
    hi = np.maximum(idx[:, None], idx[None, :])
    lo = np.minimum(idx[:, None], idx[None, :])
    pair = hi * (hi + 1) // 2 + lo  # [n,n]
    # V[p,q,r,s] = g[pair[p,r], pair[q,s]]
    V = g[pair[:, None, :, None], pair[None, :, None, :]]
    return np.ascontiguousarray(V, dtype=np.float64)


def integrals(n: int, variant: str = "S", seed: int = SEED):
    """Return (E, H, V).  variant 'S': random symmetric (throughput / sigma parity);
    variant 'D': diagonally dominant, positive semidefinite 2e part (Davidson / root parity)."""
    rng = np.random.default_rng(seed)
    npair = n * (n + 1) // 2
    if variant == "S":
        h = rng.normal(size=(n, n))
        h = (h + h.T) / 2
        g = 0.1 * rng.normal(size=(npair, npair))
        g = (g + g.T) / 2
        E = 0.5
    elif variant == "D":
        a = rng.normal(size=(n, n))
        h = np.diag(np.linspace(-1.0, 1.0, n)) + 0.05 * (a + a.T) / 2
        B = 0.05 * rng.normal(size=(3 * n, npair))
        g = B.T @ B
        E = 0.0
    else:
        raise ValueError(f"unknown synthetic integral variant {variant!r}")
    return E, np.ascontiguousarray(h), _pair_expand(g, n)


def four_fold_integrals(n: int, seed: int = SEED):
    """Random V with only <pq|rs> = <qp|sr> = <rs|pq> symmetry (no real-orbital 8-fold symmetry):
    the Harrison-Zarrabian path of the reference is valid for these, Knowles-Handy is not
    (SURVEY.md §7 hard part 8; tests/determinant/test_slater_rules.py:23-47 of the reference)."""
    rng = np.random.default_rng(seed)
    h = rng.normal(size=(n, n))
    h = (h + h.T) / 2
    v = 0.1 * rng.normal(size=(n, n, n, n))
    v = v + v.transpose(1, 0, 3, 2)
    v = v + v.transpose(2, 3, 0, 1)
    return 0.25, np.ascontiguousarray(h), np.ascontiguousarray(v)


def trial_vector(ndet: int, seed: int = SEED + 1, nvec: int = 1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(nvec, ndet))
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    return c[0] if nvec == 1 else c
