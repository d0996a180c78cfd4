"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT (only tests/, smoke() and bench.py's checker legs import this).

Size-independent parity check of a sigma vector: for a seeded sample of determinants I the exact
sigma_I = sum_J <I|H|J> C_J is rebuilt on the host from Slater rules
(forte2/determinant/slater_rules.cc:86-245, restated in oracle/sigma_port.c: port_connections) over the
determinants J of the CI space that H connects to I, with the coefficients C_J fetched from wherever the trial
vector lives (one GPU, or the row slabs of a sharded build).  Addresses come from oracle/strings_port.Layout (the
restated string order, pinned against the reference); nothing of the product is involved.  It is how the sizes
no full oracle sigma can reach (CAS(16,16), CAS(18,18), 3 GAS / 20 orbitals: BASELINE.json configs 3-5) are
compared with the reference's definition of H."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import checker, strings_port

_LAYOUTS = {}


def layout(space):
    key = repr(space)
    if key not in _LAYOUTS:
        _LAYOUTS[key] = strings_port.Layout(*space)
    return _LAYOUTS[key]


def connections(n, E, H, V, a_words, b_words):
    """per sample: (alpha words, beta words, <I|H|J>) of every connected determinant, arrays of shape (nsamp, per)"""
    L = checker.port_lib()
    if not hasattr(L, "_conn_ready"):
        up, dp = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
        L.port_connections.restype = C.c_int64
        L.port_connections.argtypes = [C.c_int, C.c_double, dp, dp, C.c_int64, up, up, up, up, dp]
        L._conn_ready = True
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    a = np.ascontiguousarray(a_words, dtype=np.uint64)
    b = np.ascontiguousarray(b_words, dtype=np.uint64)
    up, dp = C.POINTER(C.c_uint64), C.POINTER(C.c_double)
    args = (int(n), float(E), H.ctypes.data_as(dp), V.ctypes.data_as(dp), len(a), a.ctypes.data_as(up),
            b.ctypes.data_as(up))
    per = L.port_connections(*args, None, None, None)
    oa = np.zeros((len(a), per), dtype=np.uint64)
    ob = np.zeros((len(a), per), dtype=np.uint64)
    oh = np.zeros((len(a), per))
    got = L.port_connections(*args, oa.ctypes.data_as(up), ob.ctypes.data_as(up), oh.ctypes.data_as(dp))
    assert got == per, "samples with different electron counts"
    return oa, ob, oh


def sample_addresses(ndet, nsamp, seed=20261017, lo=0, hi=None):
    hi = ndet if hi is None else hi
    if hi <= lo:
        return np.zeros(0, dtype=np.int64)
    rng = np.random.default_rng(seed)
    return np.unique(rng.integers(lo, hi, size=min(nsamp, hi - lo), dtype=np.int64))


def plan(space, E, H, V, addresses, chunk=256):
    """For the sampled CI-vector addresses: yields (sample slice, J addresses (nsamp_chunk, per) with -1 outside the
    space, <I|H|J>).  Chunked so that the host arrays stay small at CAS(18,18) (9 316 connections per sample)."""
    lay = layout(space)
    addresses = np.asarray(addresses, dtype=np.int64)
    for s0 in range(0, len(addresses), chunk):
        sl = slice(s0, min(s0 + chunk, len(addresses)))
        wa, wb = lay.determinant(addresses[sl])
        ja, jb, h = connections(lay.norb, E, H, V, wa, wb)
        adr = lay.address(ja.reshape(-1), jb.reshape(-1)).reshape(ja.shape)
        assert np.array_equal(adr[:, 0], addresses[sl]), "address round trip of the sampled determinants"
        yield sl, adr, h


def sigma_at(space, E, H, V, addresses, fetch, chunk=256):
    """sigma_I for the sampled addresses; fetch(int64 address array) -> coefficients C at those addresses"""
    addresses = np.asarray(addresses, dtype=np.int64)
    out = np.zeros(len(addresses))
    for sl, adr, h in plan(space, E, H, V, addresses, chunk):
        inside = adr >= 0
        vals = np.zeros(adr.shape)
        vals[inside] = np.asarray(fetch(adr[inside]), dtype=np.float64)
        out[sl] = np.einsum("ij,ij->i", h, vals)
    return out


def torch_fetch(c):
    """fetch for a trial vector held as one torch tensor (any device)"""
    import torch

    def fetch(idx):
        return c[torch.from_numpy(idx).to(c.device)].cpu().numpy()
    return fetch


def sampled_rel_error(space, E, H, V, c, sigma, nsamp=2000, seed=20261017):
    """relative L2 error of sigma over a seeded sample of determinants: ||sigma_I - sum_J H_IJ C_J|| / ||...||;
    c and sigma are torch tensors (or numpy arrays) of the whole vector"""
    lay = layout(space)
    addresses = sample_addresses(lay.ndet, nsamp, seed)
    if isinstance(c, np.ndarray):
        want = sigma_at(space, E, H, V, addresses, lambda idx: c[idx])
        got = sigma[addresses]
    else:
        import torch
        want = sigma_at(space, E, H, V, addresses, torch_fetch(c))
        got = sigma[torch.from_numpy(addresses).to(sigma.device)].cpu().numpy()
    den = np.linalg.norm(want)
    return float(np.linalg.norm(got - want) / (den if den > 0 else 1.0)), len(addresses)


def _owner_lookup(pieces, addr):
    """local index of every CI-vector address in a rank's row slabs, -1 where the rank does not own it.
    pieces: (first local index, first address, length), ascending in both (forte2_b200.dist.Layout.pieces)."""
    if not pieces:
        return np.full(addr.shape, -1, dtype=np.int64)
    g0 = np.array([p[1] for p in pieces], dtype=np.int64)
    ln = np.array([p[2] for p in pieces], dtype=np.int64)
    l0 = np.array([p[0] for p in pieces], dtype=np.int64)
    k = np.searchsorted(g0, addr, side="right") - 1
    kk = np.maximum(k, 0)
    ok = (k >= 0) & (addr < g0[kk] + ln[kk])
    return np.where(ok, addr - g0[kk] + l0[kk], -1)


def sharded_rel_error(space, E, H, V, layouts, rank, c_loc, sigma_loc, nsamp_per_rank=256, seed=20261017, group=None):
    """The sampled check for a vector sharded by alpha rows over a torch.distributed group (forte2_b200/dist.py):
    every rank samples determinants of its own slab, lists the addresses J it needs, the requests are all-gathered,
    every rank answers for the determinants it owns, and one all-reduce returns the coefficients.  Returns
    (relative L2 error over all ranks' samples, total samples); identical on every rank."""
    import torch
    import torch.distributed as dist
    me = layouts[rank]
    world = len(layouts)
    dev = c_loc.device
    rng = np.random.default_rng(seed + 7919 * rank)
    loc = np.unique(rng.integers(0, me.local_size, size=min(nsamp_per_rank, me.local_size), dtype=np.int64))
    addresses = me.global_of_local(loc)
    chunks = list(plan(space, E, H, V, addresses))
    need = np.concatenate([adr[adr >= 0] for _, adr, _ in chunks]) if chunks else np.zeros(0, dtype=np.int64)
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    counts[rank] = len(need)
    dist.all_reduce(counts, group=group)
    nmax = int(counts.max().item())
    req = torch.full((nmax,), -1, dtype=torch.int64, device=dev)
    req[: len(need)] = torch.from_numpy(need).to(dev)
    allreq = torch.empty(world * nmax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allreq, req, group=group)
    flat = allreq.cpu().numpy()
    mine = _owner_lookup(me.pieces, flat)
    vals = torch.zeros(world * nmax, dtype=torch.float64, device=dev)
    own = np.nonzero(mine >= 0)[0]
    vals[torch.from_numpy(own).to(dev)] = c_loc[torch.from_numpy(mine[own]).to(dev)]
    dist.all_reduce(vals, group=group)
    got_c = vals[rank * nmax: rank * nmax + len(need)].cpu().numpy()
    want = np.zeros(len(addresses))
    pos = 0
    for sl, adr, h in chunks:
        inside = adr >= 0
        v = np.zeros(adr.shape)
        n_in = int(inside.sum())
        v[inside] = got_c[pos:pos + n_in]
        pos += n_in
        want[sl] = np.einsum("ij,ij->i", h, v)
    got = sigma_loc[torch.from_numpy(loc).to(dev)].cpu().numpy()
    acc = torch.tensor([float(np.sum((got - want) ** 2)), float(np.sum(want ** 2)), float(len(addresses))],
                       dtype=torch.float64, device=dev)
    dist.all_reduce(acc, group=group)
    num, den, n = (float(x) for x in acc.cpu())
    return (num ** 0.5) / (den ** 0.5 if den > 0 else 1.0), int(n)
