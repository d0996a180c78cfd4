#!/usr/bin/env python
"""End-to-end time of CISigmaBuilder.Hamiltonian(numpy, numpy) at CAS(n,n) (host path, f2_sigma_apply_host)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    E, H, V = synthetic.integrals(n, "S")
    s = f2.CIStrings(n // 2, n // 2, 0, [[0] * n], [], [])
    b = f2.CISigmaBuilder(s, E, H, V, 0)
    b.set_memory(96 * 1024)
    c = synthetic.trial_vector(s.ndet)
    out = np.empty_like(c)
    for _ in range(2):
        b.Hamiltonian(c, out)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter()
        b.Hamiltonian(c, out)
        ts.append((time.perf_counter() - t0) * 1e3)
    cd = torch.from_numpy(c).cuda()
    sd = torch.empty_like(cd)
    b.Hamiltonian(cd, sd)
    torch.cuda.synchronize()
    err = float(np.linalg.norm(out - sd.cpu().numpy()) / np.linalg.norm(out))
    print({"n": n, "e2e_ms": [round(t, 2) for t in ts], "resident_ms": round(b.last_build_ms()[4], 2), "rel_diff_vs_device_path": err,
           "env": {k: v for k, v in os.environ.items() if k.startswith("F2_")}}, flush=True)


if __name__ == "__main__":
    main()
