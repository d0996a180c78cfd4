#!/usr/bin/env python
"""Time sigma builds of a named non-CAS space on one GPU (CUDA events), with the Hermiticity invariant.
    python scripts/space_time.py gas20twin [steps]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def main():
    import torch

    import forte2_b200 as f2
    from forte2_b200 import synthetic
    from sharded_run import WORKLOADS
    name = sys.argv[1] if len(sys.argv) > 1 else "gas20twin"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    space = WORKLOADS[name]
    n = sum(len(g) for g in space[3])
    t0 = time.perf_counter()
    E, H, V = synthetic.integrals(n, "S")
    strings = f2.CIStrings(*space)
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    b.set_memory(32 * 1024)
    setup = time.perf_counter() - t0
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(strings.ndet, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(strings.ndet, dtype=torch.float64, device="cuda", generator=g)
    hx, hy = torch.empty_like(x), torch.empty_like(y)
    b.Hamiltonian(x, hx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        b.Hamiltonian(x, hx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    last = b.last_build_ms()
    b.Hamiltonian(y, hy)
    def dot(a, b, step=1 << 30):  # torch.dot stops at 2^31 - 1 elements
        return sum(float(a[i:i + step] @ b[i:i + step]) for i in range(0, a.numel(), step))

    herm = abs(dot(y, hx) - dot(x, hy)) / (dot(hx, hx) ** 0.5 * dot(y, y) ** 0.5)
    shape = b.t3_shape()
    print(json.dumps({"space": name, "ndet": strings.ndet, "nblocks": strings.nblocks, "setup_s": round(setup, 2),
                      "sigma_ms": ms, "term_ms": {"H0": last[0], "same_spin_alpha": last[1], "same_spin_beta": last[2],
                                                  "opposite_spin": last[3], "total": last[4]},
                      "t3_form": shape["form"], "t3_useful_tflops": 2.0 * shape["M"] * shape["k_cols"] / (last[3] * 1e-3) / 1e12,
                      "hermiticity_rel": herm, "env": {k: v for k, v in os.environ.items() if k.startswith("F2_")}}))


if __name__ == "__main__":
    main()
