#!/usr/bin/env python
"""Kernel times of ONE rank's share of the sharded (rows-form) sigma build, emulated on one GPU: rank 0 of `world` runs
f2_sigma_apply_rows on its row slab and column slab (no exchange), per-term CUDA-event times of the last build.
    python scripts/rows_time.py cas16 8 [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    from forte2_b200.dist import CudaEngine, Layout
    from sharded_run import WORKLOADS
    name = sys.argv[1] if len(sys.argv) > 1 else "cas16"
    world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    space = WORKLOADS[name]
    n = sum(len(g) for g in space[3])
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    b.set_memory(16384)
    eng = CudaEngine(b)
    ok, slot = eng.rows_form(0, world)
    assert ok
    L = Layout(strings.blocks(), 0, world)
    g = torch.Generator(device="cuda").manual_seed(5)
    c_loc = torch.randn(L.local_size, dtype=torch.float64, device="cuda", generator=g)
    xcols = torch.randn(max(L.exchange_size, 1), dtype=torch.float64, device="cuda", generator=g)
    sigma_loc = torch.empty_like(c_loc)
    ycols = torch.empty_like(xcols)
    scatter = torch.empty(world * slot, dtype=torch.float64, device="cuda")
    for _ in range(steps):
        eng.apply_rows(c_loc, xcols, sigma_loc, ycols, scatter, 0, world)
    torch.cuda.synchronize()
    ms = b.last_build_ms()
    print(json.dumps({"space": name, "world": world, "rank": 0, "local": L.local_size, "exchange": L.exchange_size,
                      "term_ms": dict(zip(["H0", "same_spin_alpha_cols", "same_spin_beta_rows", "opposite_spin", "total"],
                                          [round(x, 3) for x in ms])),
                      "env": {k: v for k, v in os.environ.items() if k.startswith("F2_")}}), flush=True)


if __name__ == "__main__":
    main()
