#!/usr/bin/env python
"""A few Davidson-Liu iterations of a 3-root solve (CAS(16,16) by default), for timing the solver's own passes:
run under `ncu --metrics gpu__time_duration.sum` for the launch list, or alone for the per-iteration summary."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    from forte2_b200.ci_solver import CISolver
    from forte2_b200.davidson import DavidsonLiuParams
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    mult = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    E, H, V = synthetic.integrals(n, "D")
    strings = f2.CIStrings(n // 2, n // 2, 0, [[0] * n], [], [])
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    b.set_memory(96 * 1024)
    s = CISolver(strings, E, H, V, nroot=3, davidson_liu_params=DavidsonLiuParams(maxiter=iters), log_level=0, builder=b,
                 multiplicity=mult or None)
    s.run()
    ds = s.eigensolver
    print(json.dumps({"n": n, "iter_ms": [round(t * 1e3, 2) for t in ds.iter_times],
                      "iter_gbytes": [round(x / 1e9, 1) for x in ds.iter_bytes],
                      "sigma_builds": [h[5] for h in ds.history], "basis": [h[4] for h in ds.history],
                      "sigma_ms": b.last_build_ms()[4], "E": [float(x) for x in s.E]}))


if __name__ == "__main__":
    main()
