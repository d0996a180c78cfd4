#!/usr/bin/env python
"""Setup time of a space: string enumeration, replacement lists (GPU kernels or host builder), builder creation.
usage: setup_time.py {cas16|cas18|gas20|gas20twin} [host]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CFGS = {"gas20": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [8, 0, 0], [12, 16, 4]),
        "cas18": (9, 9, 0, [[0] * 18], [], []),
        "cas16": (8, 8, 0, [[0] * 16], [], []),
        "gas20twin": (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])}


def main():
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    name = sys.argv[1]
    host = len(sys.argv) > 2 and sys.argv[2] == "host"
    torch.zeros(1, device="cuda")  # context
    torch.cuda.synchronize()
    space = CFGS[name]
    n = sum(len(s) for s in space[3])
    E, H, V = synthetic.integrals(n, "S")
    t0 = time.perf_counter()
    s = f2.CIStrings(*space)
    t1 = time.perf_counter()
    s.build_lists(None if host else 0)
    t2 = time.perf_counter()
    out = {"space": name, "ndet": s.ndet, "nas": s.nas, "lists": "host" if host else "device",
           "strings_s": round(t1 - t0, 3), "lists_s": round(t2 - t1, 3)}
    if s.ndet * 16 < 150e9:
        b = f2.CISigmaBuilder(s, E, H, V, 0)
        torch.cuda.synchronize()
        out["builder_s"] = round(time.perf_counter() - t2, 3)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
