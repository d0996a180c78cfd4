#!/usr/bin/env python
"""Sharded sigma build + Davidson-Liu iterations on N GPUs of one box (torchrun), for the active spaces
that need more than one GPU (BASELINE.json configs[3]: CAS(18e,18o), 2 363 904 400 determinants).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 scripts/sharded_run.py --workload cas18 --steps 3 --davidson-iters 3

Prints one JSON line (rank 0): time per sigma build (CUDA events, max over ranks), per-term kernel times
of rank 0, the Hermiticity invariant <x|H y> = <y|H x> and the time
per Davidson iteration of a 1-root solve.  Oracle-free: the reference cannot run this size.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def _cas(n):
    return (n // 2, n // 2, 0, [[0] * n], [], [])


# name -> CIStrings arguments.  gas20 is BASELINE.json configs[4] as made concrete in SURVEY.md section 8d
# (1 865 504 274 determinants, 225 blocks); gas20twin its 1-GPU validation size (22 065 484 determinants).
WORKLOADS = {"cas12": _cas(12), "cas14": _cas(14), "cas16": _cas(16), "cas18": _cas(18),
             "gas20": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [8, 0, 0], [12, 16, 4]),
             "gas20twin": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2]),
             # CAS with point-group symmetry: 8 irreps, two orbitals each (the usual forte2 situation)
             "cas14_d2h": (7, 7, 0, [[i % 8 for i in range(14)]], [], []),
             "cas16_d2h": (8, 8, 0, [[i % 8 for i in range(16)]], [], []),
             "cas16_c2v": (8, 8, 0, [[i % 4 for i in range(16)]], [], [])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cas18", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--davidson-iters", type=int, default=3)
    ap.add_argument("--nroot", type=int, default=1)
    ap.add_argument("--form", default="auto")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    import forte2_b200 as f2
    from forte2_b200 import synthetic
    from forte2_b200.ci_solver import ShardedCISolver
    from forte2_b200.davidson import DavidsonLiuParams
    from forte2_b200.dist import ShardedSigma

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist.init_process_group("nccl", device_id=dev)

    space = WORKLOADS[args.workload]
    na, nb = space[0], space[1]
    n = sum(len(g) for g in space[3])
    t0 = time.perf_counter()
    E, H, V = synthetic.integrals(n, "D")  # diagonally dominant variant: the Davidson leg converges
    strings = f2.CIStrings(*space)
    builder = f2.CISigmaBuilder(strings, E, H, V, 0, device=local_rank)
    builder.set_memory(64 * 1024)
    sh = ShardedSigma(builder, rank, world, form=args.form)
    setup_s = time.perf_counter() - t0

    gen = torch.Generator(device=dev)
    gen.manual_seed(20261018 + rank)
    x = torch.randn(sh.local_size, dtype=torch.float64, device=dev, generator=gen)
    y = torch.randn(sh.local_size, dtype=torch.float64, device=dev, generator=gen)
    hx, hy = torch.empty_like(x), torch.empty_like(y)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        sh.Hamiltonian(x, hx)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        sh.Hamiltonian(x, hx)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    last = builder.last_build_ms()

    # Hermiticity: <y|H x> == <x|H y>
    sh.Hamiltonian(y, hy)
    dots = torch.stack([torch.dot(y, hx), torch.dot(x, hy), torch.dot(hx, hx), torch.dot(hy, hy),
                        torch.dot(x, x), torch.dot(y, y)])
    dist.all_reduce(dots, op=dist.ReduceOp.SUM)
    yHx, xHy, nhx, nhy, nx, ny = [float(v) for v in dots.cpu()]
    herm = abs(yHx - xHy) / ((nhx ** 0.5) * (ny ** 0.5))
    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    del x, y, hx, hy
    torch.cuda.empty_cache()

    davidson = None
    if args.davidson_iters > 0:
        params = DavidsonLiuParams(maxiter=args.davidson_iters)
        solver = ShardedCISolver(strings, E, H, V, nroot=args.nroot, davidson_liu_params=params, log_level=0,
                                 device=local_rank, builder=builder, form=args.form)
        solver.sharded = sh
        t0 = time.perf_counter()
        try:
            solver.run()
        except Exception as e:  # maxiter reached is not an error for a timing run
            if "converge" not in str(e).lower():
                raise
        ds = solver.eigensolver
        davidson = {"nroot": args.nroot, "iterations_run": len(ds.iter_times),
                    "iter_ms_all": [round(t * 1e3, 2) for t in ds.iter_times],
                    "sigma_builds_per_iter": [h[5] for h in ds.history],
                    "energies_last": [float(v) for v in ds.lam[: args.nroot]],
                    "max_residual_norm_last": float(ds.history[-1][3]) if ds.history else None,
                    "wall_s": time.perf_counter() - t0}
        mem_gb = max(mem_gb, torch.cuda.max_memory_allocated(dev) / 1e9)

    if rank == 0:
        shape = builder.t3_shape()
        ndet = strings.ndet
        useful = 2.0 * shape["M"] * shape["k_cols"] / world  # k_cols: contracted (k, column) pairs of the full build
        line = {
            "what": "sharded sigma build", "workload": args.workload, "ndet": ndet, "nblocks": strings.nblocks,
            "n_gpus": world,
            "form": sh.form, "setup_s": round(setup_s, 2),
            "sigma_ms": float(ms.item()), "sigma_builds_per_s": 1e3 / float(ms.item()),
            "term_ms_rank0": {"H0": last[0], "same_spin_alpha": last[1], "same_spin_beta": last[2],
                              "opposite_spin": last[3], "total": last[4]},
            "exchange_ms_rank0": float(ms.item()) - last[4],
            "opposite_spin_useful_tflops_per_gpu": useful / (last[3] * 1e-3) / 1e12 if last[3] else None,
            "hermiticity_rel": herm, "local_vector_gb": sh.local_size * 8 / 1e9,
            "peak_mem_gb_rank0": mem_gb, "davidson": davidson,
        }
        print(json.dumps(line), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
