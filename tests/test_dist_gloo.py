"""world_size = 2 and 3 over gloo on CPU: the row partition and both communication forms of
forte2_b200/dist.py (full: all-gather + column-owned alpha same-spin exchange; rows: all-to-all of the
column slab, exchange back, reduce-scatter of the opposite-spin term), with the oracle port standing in
for the GPU's local work.  The GPU version of the same check (NCCL, the real C ABI) is
tests/test_dist_gpu.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cases import SMALL_SPACES
from forte2_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class PortEngine:
    """CPU stand-in for CudaEngine: same contract as f2_sigma_apply_part, computed with oracle/sigma_port.c
    (scalar + beta + alpha-beta terms for the owned rows; alpha same-spin term for the owned columns)."""

    def __init__(self, dets, ints, layouts):
        self.dets, self.ints, self.layouts = dets, ints, layouts

    def apply_part(self, c_full, sigma_loc, yexch, rank, world):
        from oracle import checker
        E, H, V = self.ints
        c = c_full.numpy()
        rows_part = checker.port_sigma(self.dets, E, H, V, c, checker.PART_SCALAR | checker.PART_BETA |
                                       checker.PART_ALPHA_BETA)
        alpha_part = checker.port_sigma(self.dets, E, H, V, c, checker.PART_ALPHA)
        lay = self.layouts[rank]
        out = sigma_loc.numpy()
        y = yexch.numpy()
        for (off, nA, nB), (lo, n, loc), (cl, nc, yoff) in zip(lay.blocks, lay.rows, lay.cols):
            if nA * nB == 0:
                continue
            out[loc:loc + n * nB] = rows_part[off + lo * nB: off + (lo + n) * nB]
            if world > 1:
                y[yoff:yoff + nA * nc] = alpha_part[off:off + nA * nB].reshape(nA, nB)[:, cl:cl + nc].reshape(-1)
            else:
                out[loc:loc + n * nB] += alpha_part[off + lo * nB: off + (lo + n) * nB]


    # ---- rows form: same contract as f2_sigma_rows_form / f2_sigma_apply_rows ----
    def rows_form(self, rank, world):
        return True, max(L.local_size for L in self.layouts)

    def apply_rows(self, c_loc, xcols, sigma_loc, ycols, scatter, rank, world, phases=3):
        """phases as in f2_sigma_apply_rows: 1 opposite-spin term (local rows only), 2 = 4 | 8, 4 H0 + alpha same-spin
        term (from xcols), 8 beta same-spin term (local rows)"""
        from oracle import checker
        E, H, V = self.ints
        lay = self.layouts[rank]
        rows_only = np.zeros(lay.ndet)
        cols_only = np.zeros(lay.ndet)
        x = xcols.numpy()
        for (off, nA, nB), (lo, n, loc), (cl, nc, yoff) in zip(lay.blocks, lay.rows, lay.cols):
            if nA * nB == 0:
                continue
            rows_only[off + lo * nB: off + (lo + n) * nB] = c_loc.numpy()[loc:loc + n * nB]
            cols_only[off:off + nA * nB].reshape(nA, nB)[:, cl:cl + nc] = x[yoff:yoff + nA * nc].reshape(nA, nc)
        if phases & (2 | 4):
            local = checker.port_sigma(self.dets, E, H, V, rows_only, checker.PART_SCALAR)
            alpha = checker.port_sigma(self.dets, E, H, V, cols_only, checker.PART_ALPHA)
            y = ycols.numpy()
            for (off, nA, nB), (lo, n, loc), (cl, nc, yoff) in zip(lay.blocks, lay.rows, lay.cols):
                if nA * nB == 0:
                    continue
                sigma_loc.numpy()[loc:loc + n * nB] = local[off + lo * nB: off + (lo + n) * nB]
                y[yoff:yoff + nA * nc] = alpha[off:off + nA * nB].reshape(nA, nB)[:, cl:cl + nc].reshape(-1)
        if phases & (2 | 8):
            beta = checker.port_sigma(self.dets, E, H, V, rows_only, checker.PART_BETA)
            for (off, nA, nB), (lo, n, loc), (cl, nc, yoff) in zip(lay.blocks, lay.rows, lay.cols):
                if nA * nB > 0:
                    sigma_loc.numpy()[loc:loc + n * nB] += beta[off + lo * nB: off + (lo + n) * nB]
        if phases & 1:
            t3 = checker.port_sigma(self.dets, E, H, V, rows_only, checker.PART_ALPHA_BETA)
            _, slot = self.rows_form(rank, world)
            out = scatter.numpy()
            out[:] = 0.0
            for q, L in enumerate(self.layouts):
                for (off, nA, nB), (lo_q, n_q, loc_q) in zip(L.blocks, L.rows):
                    if nA * nB > 0:
                        out[q * slot + loc_q: q * slot + loc_q + n_q * nB] = t3[off + lo_q * nB: off + (lo_q + n_q) * nB]


def _worker(rank, world, port, name, form, result_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forte2_b200 as f2
        from forte2_b200.dist import ShardedSigma
        from oracle import checker
        space = SMALL_SPACES[name]
        n = sum(len(s) for s in space[3])
        ints = synthetic.integrals(n, "S")
        strings = f2.CIStrings(*space)
        dets = strings.make_determinants()
        from forte2_b200.dist import Layout
        engine = PortEngine(dets, ints, [Layout(strings.blocks(), q, world) for q in range(world)])
        sh = ShardedSigma(rank=rank, world=world, engine=engine, blocks=strings.blocks(),
                          device=torch.device("cpu"), form=form)
        assert sh.form == form
        c_full = torch.from_numpy(synthetic.trial_vector(strings.ndet))
        c_loc = sh.scatter_full(c_full)
        sigma_loc = torch.zeros_like(c_loc)
        sh.Hamiltonian(c_loc, sigma_loc)
        got = sh.all_gather_full(sigma_loc).clone().numpy()
        want = checker.port_sigma(dets, *ints, c_full.numpy())
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert sum(lay.local_size for lay in sh.layouts) == strings.ndet
        # the sampled checker in its sharded form (oracle/sampled.py: what bench.py runs at CAS(18,18) on 8 GPUs)
        from oracle import sampled
        serr, ns = sampled.sharded_rel_error(space, *ints, sh.layouts, rank, c_loc, sigma_loc, nsamp_per_rank=48)
        assert ns > 0 and serr <= 1e-13, (serr, ns)
        bad = sigma_loc.clone()
        bad[: max(1, bad.numel() // 2)] *= 1.0 + 1e-8
        assert sampled.sharded_rel_error(space, *ints, sh.layouts, rank, c_loc, bad, nsamp_per_rank=48)[0] > 1e-10
        np.save(os.path.join(result_dir, f"err_{rank}.npy"), np.array([err]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("name,world,form", [("cas66", 2, "full"), ("gas343_54", 2, "full"), ("cas66_d2h", 3, "full"),
                                             ("cas75_odd", 3, "full"), ("cas66", 2, "rows"), ("cas75_odd", 3, "rows"),
                                             ("cas88", 3, "rows"), ("gas343_54", 2, "rows"), ("cas66_d2h", 3, "rows"),
                                             ("gas2_sym4", 2, "rows")])
def test_sharded_sigma_over_gloo(tmp_path, name, world, form):
    mp.spawn(_worker, args=(world, _free_port(), name, form, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err = float(np.load(tmp_path / f"err_{r}.npy")[0])
        assert err <= 1e-13, (r, err)


def _solver_worker(rank, world, port, name, form, result_dir, multiplicity=None):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import forte2_b200 as f2
        from forte2_b200.ci_solver import ShardedCISolver
        from forte2_b200.davidson import DavidsonLiuParams
        from forte2_b200.dist import Layout, ShardedSigma
        from host_dl_kernels import HostKernels
        from oracle import checker
        space = SMALL_SPACES[name]
        n = sum(len(s) for s in space[3])
        ints = synthetic.integrals(n, "D")
        strings = f2.CIStrings(*space)
        dets = strings.make_determinants()
        engine = PortEngine(dets, ints, [Layout(strings.blocks(), q, world) for q in range(world)])
        sh = ShardedSigma(rank=rank, world=world, engine=engine, blocks=strings.blocks(),
                          device=torch.device("cpu"), form=form)
        e_det = checker.port_energies(dets, *ints)
        if multiplicity is not None:
            # the spin-symmetric diagonal (f2_sigma_hdiag_spin_averaged on the GPU): mean over each orbital configuration
            conf = {}
            for k, (a, b) in enumerate(dets):
                conf.setdefault((int(a & b), int(a ^ b)), []).append(k)
            for members in conf.values():
                e_det[members] = e_det[members].mean()
        hdiag = sh.scatter_full(torch.from_numpy(e_det))
        # (residuals much below 1e-8 are out of reach by construction: corrections shorter than the reference's
        # schmidt_discard_threshold = 1e-8 are dropped, davidsonliu.py:460-463)
        solver = ShardedCISolver(strings, *ints, nroot=2,
                                 davidson_liu_params=DavidsonLiuParams(r_tol=1e-8 if multiplicity is None else 1e-7),
                                 log_level=0, _sharded=sh, _hdiag=hdiag, multiplicity=multiplicity,
                                 _kernels=HostKernels(comm=dist.group.WORLD)).run()
        assert solver.converged
        assert (multiplicity is None) == (getattr(solver.eigensolver, "_proj_sparse", None) is None)
        vec_norms = sh.all_gather_full(solver.evecs[0].contiguous()).norm()
        np.save(os.path.join(result_dir, f"roots_{rank}.npy"), np.append(np.asarray(solver.E), float(vec_norms)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,world,form", [("cas66", 2, "rows"), ("cas66_d2h", 3, "full"), ("cas66_d2h", 2, "rows")])
def test_sharded_davidson_over_gloo(tmp_path, name, world, form):
    """ShardedCISolver (guess determinants agreed through an all-gather, every Davidson vector a row
    slab, reductions all-reduced) reproduces the two lowest eigenvalues of the dense Hamiltonian."""
    from oracle import checker
    import forte2_b200 as f2
    mp.spawn(_solver_worker, args=(world, _free_port(), name, form, str(tmp_path)), nprocs=world, join=True)
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    ints = synthetic.integrals(n, "D")
    dets = f2.CIStrings(*space).make_determinants()
    eye = np.eye(len(dets))
    Hd = np.stack([checker.port_sigma(dets, *ints, eye[i]) for i in range(len(dets))])
    want = np.linalg.eigvalsh(Hd)[:2]
    for r in range(world):
        got = np.load(tmp_path / f"roots_{r}.npy")
        assert np.max(np.abs(got[:2] - want)) <= 1e-10, (r, got, want)
        assert abs(got[2] - 1.0) <= 1e-10


@pytest.mark.parametrize("name,world,multiplicity", [("cas66", 2, 1), ("cas66", 3, 3), ("cas75_odd", 2, 2)])
def test_sharded_davidson_with_spin_control_over_gloo(tmp_path, name, world, multiplicity):
    """determinant-basis spin control on sharded vectors (spin-pure guesses, wrong-spin guess combinations projected
    out with an all-reduce of the overlaps): the two lowest roots of the wanted multiplicity"""
    from oracle import checker
    import forte2_b200 as f2
    from forte2_b200.ci_solver import spin2_matrix
    mp.spawn(_solver_worker, args=(world, _free_port(), name, "rows", str(tmp_path), multiplicity), nprocs=world,
             join=True)
    space = SMALL_SPACES[name]
    n = sum(len(s) for s in space[3])
    ints = synthetic.integrals(n, "D")
    strings = f2.CIStrings(*space)
    dets = strings.make_determinants()
    eye = np.eye(len(dets))
    Hd = np.stack([checker.port_sigma(dets, *ints, eye[i]) for i in range(len(dets))])
    S = 0.5 * (multiplicity - 1)
    D = spin2_matrix(strings, np.arange(len(dets))) - S * (S + 1) * eye
    w, U = np.linalg.eigh(Hd + 1e3 * (D @ D))
    want = w[:2]  # the penalty removes every other multiplicity from the bottom of the spectrum
    for r in range(world):
        got = np.load(tmp_path / f"roots_{r}.npy")
        assert np.max(np.abs(got[:2] - want)) <= 1e-9, (r, got, want)
