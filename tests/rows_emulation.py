"""The rows form of the sharded sigma build (forte2_b200/dist.py) for every rank of a world, run one after the other on
ONE GPU with the three exchanges done by hand: what tests/test_sigma_gpu.py and tests/test_sampled_parity_gpu.py use
to exercise the sharded kernel instances at their natural sizes on the driver's single-GPU test box."""


def rows_form_sigma(torch, strings, builder, C, world, nan_fill=True):
    from forte2_b200.dist import CudaEngine, Layout
    eng = CudaEngine(builder)
    lays = [Layout(strings.blocks(), q, world) for q in range(world)]
    ok, slot = eng.rows_form(0, world)
    assert ok and slot == max(L.local_size for L in lays)
    fill = float("nan") if nan_fill else 0.0
    total = torch.zeros_like(C)
    for q, L in enumerate(lays):
        # several blocks: the vectors need no 16-byte alignment (rows of a Davidson basis with an odd length)
        skew = 1 if strings.nblocks > 1 else 0
        c_loc = torch.empty(L.local_size + 1, dtype=torch.float64, device="cuda")[skew: skew + L.local_size]
        xcols = torch.full((max(L.exchange_size, 1),), fill, dtype=torch.float64, device="cuda")
        for (off, nA, nB), (lo, nr, loc), (cl, nc, yoff) in zip(L.blocks, L.rows, L.cols):
            if nA * nB == 0:
                continue
            blk = C[off:off + nA * nB].view(nA, nB)
            c_loc[loc:loc + nr * nB] = blk[lo:lo + nr].reshape(-1)
            xcols[yoff:yoff + nA * nc] = blk[:, cl:cl + nc].reshape(-1)
        sigma_loc = torch.full((L.local_size + 1,), fill, dtype=torch.float64, device="cuda")[skew: skew + L.local_size]
        ycols = torch.full_like(xcols, fill)
        scatter = torch.full((world * slot,), fill, dtype=torch.float64, device="cuda")
        eng.apply_rows(c_loc, xcols, sigma_loc, ycols, scatter, q, world)
        for (off, nA, nB), (lo, nr, loc), (cl, nc, yoff) in zip(L.blocks, L.rows, L.cols):
            if nA * nB == 0:
                continue
            blk = total[off:off + nA * nB].view(nA, nB)
            blk[lo:lo + nr] += sigma_loc[loc:loc + nr * nB].view(nr, nB)
            blk[:, cl:cl + nc] += ycols[yoff:yoff + nA * nc].view(nA, nc)
        for r, Lr in enumerate(lays):
            for (off, nA, nB), (lo_r, n_r, loc_r) in zip(Lr.blocks, Lr.rows):
                if nA * nB > 0:
                    total[off + lo_r * nB: off + (lo_r + n_r) * nB] += scatter[r * slot + loc_r: r * slot + loc_r + n_r * nB]
        del c_loc, xcols, sigma_loc, ycols, scatter
    return total
