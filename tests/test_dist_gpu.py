"""N = 2 GPUs over NCCL (skipped with fewer): sharded sigma through the real C ABI equals the
single-GPU sigma and the reference to 1e-12 relative L2."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _die():
    """a rank that failed leaves at once: tearing the process group down would wait for a peer that is
    itself waiting in a collective, and the test would only end at its deadline"""
    import traceback
    traceback.print_exc()
    sys.stderr.flush()
    os._exit(1)


def _worker(rank, world, port, spaces, result_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import forte2_b200 as f2
        from forte2_b200 import synthetic
        from forte2_b200.dist import ShardedSigma
        errs = []
        for space in spaces:
            n = sum(len(s) for s in space[3])
            E, H, V = synthetic.integrals(n, "S")
            strings = f2.CIStrings(*space)
            builder = f2.CISigmaBuilder(strings, E, H, V, 0, device=rank)
            c_full = torch.from_numpy(synthetic.trial_vector(strings.ndet)).cuda()
            want = torch.empty_like(c_full)
            builder.Hamiltonian(c_full, want)  # single-GPU result on this rank
            forms = []
            for form in ("full", "auto"):
                sh = ShardedSigma(builder, rank, world, form=form)
                if form == "auto" and sh.form == "full":
                    continue  # the rows form does not exist for this space; "full" is already checked
                forms.append(sh.form)
                c_loc = sh.scatter_full(c_full)
                sigma_loc = torch.full_like(c_loc, float("nan"))
                sh.Hamiltonian(c_loc, sigma_loc)
                sh.Hamiltonian(c_loc, sigma_loc)  # twice: no dependence on leftovers
                got = sh.all_gather_full(sigma_loc).clone()
                errs.append(float((got - want).norm() / want.norm()))
                # the shard from and to host memory (pageable numpy, then pinned tensors)
                c_np = c_loc.cpu().numpy()
                s_np = np.full_like(c_np, np.nan)
                sh.Hamiltonian_host(c_np, s_np)
                c_pin = torch.empty(c_loc.numel(), dtype=torch.float64, pin_memory=True).copy_(c_loc)
                s_pin = torch.full((c_loc.numel(),), float("nan"), dtype=torch.float64).pin_memory()
                sh.Hamiltonian_host(c_pin, s_pin)
                assert np.linalg.norm(s_np - s_pin.numpy()) <= 1e-12 * np.linalg.norm(s_np)  # reductions commute, not bitwise
                errs.append(float((sh.all_gather_full(s_pin.cuda()) - want).norm() / want.norm()))
            if "rows" in forms:
                # the same build with the collectives issued by the C library (f2_dist_init / f2_dist_apply)
                from forte2_b200.dist import LibraryShardedSigma
                lib = LibraryShardedSigma(builder, rank, world)
                c_loc = sh.scatter_full(c_full)
                sigma_loc = torch.full_like(c_loc, float("nan"))
                lib.Hamiltonian(c_loc, sigma_loc)
                lib.Hamiltonian(c_loc, sigma_loc)
                torch.cuda.synchronize()
                got = sh.all_gather_full(sigma_loc).clone()
                errs.append(float((got - want).norm() / want.norm()))
                one = torch.ones(3, dtype=torch.float64, device="cuda")
                assert float(lib.allreduce_(one).sum()) == 3.0 * world
                lib.close()
            if n >= 8 and len(space[3]) == 1 and strings.nbs % 2 == 0 and strings.nas % 2 == 0:
                assert "rows" in forms, forms
            if strings.nblocks > 1:  # symmetry / GAS blocks: the block-aware gather kernel serves the rows form
                assert "rows" in forms, forms
        np.save(os.path.join(result_dir, f"err_{rank}.npy"), np.array(errs))
        dist.destroy_process_group()
    except BaseException:
        _die()


def _spawn(fn, args, nprocs, timeout=150):
    """mp.spawn with a deadline: a collective mismatch would otherwise hang the GPU box until NCCL's
    watchdog fires"""
    import time
    import torch.multiprocessing as mp
    ctx = mp.spawn(fn, args=args, nprocs=nprocs, join=False)
    deadline = time.time() + timeout
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p in ctx.processes:
                if p.is_alive():
                    p.kill()
            pytest.fail(f"{fn.__name__}: {nprocs} ranks did not finish within {timeout} s")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_sigma_two_gpus(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    spaces = [(3, 3, 0, [[0] * 6], [], []), (5, 4, 0, [[0] * 3, [0] * 4, [0] * 3], [4, 0, 0], [6, 8, 2]),
              (3, 3, 0, [[0, 5, 2, 3, 6, 7]], [], []), (5, 5, 0, [[0] * 10], [], []), (4, 3, 0, [[0] * 7], [], []),
              (6, 6, 0, [[0] * 12], [], []), (5, 5, 0, [[0] * 3, [0] * 4, [0] * 3], [4, 0, 0], [6, 8, 2]),
              (4, 4, 0, [[0, 1, 2, 3] * 2], [], []), (4, 3, 2, [[0, 1, 2], [3, 0, 1, 2]], [2, 0], [6, 7])]
    _spawn(_worker, (2, _free_port(), spaces, str(tmp_path)), 2)
    for r in range(2):
        errs = np.load(tmp_path / f"err_{r}.npy")
        assert np.all(errs <= 1e-12), (r, errs)


def _solver_worker(rank, world, port, result_dir):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import forte2_b200 as f2
        from forte2_b200 import synthetic
        from forte2_b200.ci_solver import CISolver, ShardedCISolver
        from forte2_b200.davidson import DavidsonLiuParams
        out = []
        for space, form in [((5, 5, 0, [[0] * 10], [], []), "auto"), ((5, 5, 0, [[0] * 10], [], []), "full"),
                            ((5, 4, 0, [[0] * 3, [0] * 4, [0] * 3], [4, 0, 0], [6, 8, 2]), "auto")]:
            n = sum(len(s) for s in space[3])
            E, H, V = synthetic.integrals(n, "D")
            strings = f2.CIStrings(*space)
            params = DavidsonLiuParams(e_tol=1e-12, r_tol=1e-8)
            single = CISolver(strings, E, H, V, nroot=3, davidson_liu_params=params, log_level=0, device=rank).run()
            sharded = ShardedCISolver(strings, E, H, V, nroot=3, davidson_liu_params=params, log_level=0,
                                      device=rank, form=form).run()
            out.append(np.concatenate([np.asarray(single.E), np.asarray(sharded.E)]))
            # RDMs of a root of the sharded solve (full copies, work split over the ranks, all-reduce) against the single-GPU
            # ones; the roots can differ by a sign, the RDMs of a state with itself cannot
            for kind in ("make_sf_1rdm", "make_sf_2rdm"):
                a, b2 = getattr(single, kind)(1), getattr(sharded, kind)(1)
                assert np.max(np.abs(a - b2)) <= 1e-7, (kind, float(np.max(np.abs(a - b2))))
        np.save(os.path.join(result_dir, f"roots_{rank}.npy"), np.array(out))
        dist.destroy_process_group()
    except BaseException:
        _die()


def test_sharded_davidson_roots_two_gpus(tmp_path):
    """3-root Davidson-Liu with every vector sharded over 2 GPUs == the single-GPU roots to 1e-10 Eh
    (BASELINE.json), in both communication forms and on a GAS space."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _spawn(_solver_worker, (2, _free_port(), str(tmp_path)), 2)
    for r in range(2):
        roots = np.load(tmp_path / f"roots_{r}.npy")
        for row in roots:
            assert np.max(np.abs(row[:3] - row[3:])) <= 1e-10, (r, row)


def test_one_strings_object_two_devices():
    """one CIStrings shared by builders on cuda:0 and cuda:1 in the same process (each device keeps its own copy of the
    lists for as long as the strings object lives): alternating builds give the same sigma on both"""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, ROOT)
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    space = (4, 4, 0, [[0] * 8], [], [])
    E, H, V = synthetic.integrals(8, "S")
    strings = f2.CIStrings(*space)
    b0 = f2.CISigmaBuilder(strings, E, H, V, 0, device=0)
    b1 = f2.CISigmaBuilder(strings, E, H, V, 0, device=1)
    c = synthetic.trial_vector(strings.ndet)
    out = []
    for rep in range(2):
        for dev, b in ((0, b0), (1, b1)):
            with torch.cuda.device(dev):
                cd = torch.from_numpy(c).to(f"cuda:{dev}")
                sd = torch.empty_like(cd)
                b.Hamiltonian(cd, sd)
                torch.cuda.synchronize(dev)
                out.append(sd.cpu().numpy())
                g1 = b.a_1rdm(cd, cd)
                assert abs(np.trace(g1) - 4.0 * float(c @ c)) <= 1e-10
    del b0  # the first builder goes away; the second one's lists stay valid
    with torch.cuda.device(1):
        cd = torch.from_numpy(c).to("cuda:1")
        sd = torch.empty_like(cd)
        b1.Hamiltonian(cd, sd)
        out.append(sd.cpu().numpy())
    for o in out[1:]:
        assert np.linalg.norm(o - out[0]) <= 1e-13 * np.linalg.norm(out[0])
