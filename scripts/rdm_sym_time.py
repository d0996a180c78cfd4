#!/usr/bin/env python
"""Opposite-spin 2-RDM (and the other 1-/2-body blocks) of CAS(n,n) with D2h-like orbital irreps (8 irreps): the
irrep-blocked D-matrix path against the unblocked one (F2_RDM_NO_IRREP_BLOCKS).  usage: rdm_sym_time.py n"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    E, H, V = synthetic.integrals(n, "S")
    s = f2.CIStrings(n // 2, n // 2, 0, [[i % 8 for i in range(n)]], [], [])
    b = f2.CISigmaBuilder(s, E, H, V, 0)
    c = torch.from_numpy(synthetic.trial_vector(s.ndet)).cuda()
    out = {}
    for kind in ("a_1rdm", "b_1rdm", "aa_2rdm", "bb_2rdm", "ab_2rdm"):
        getattr(b, kind)(c, c)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = getattr(b, kind)(c, c)
        out[kind] = round((time.perf_counter() - t0) * 1e3, 2)
        if kind == "ab_2rdm":
            out["ab_trace"] = float(sum(g[u, v, u, v] for u in range(n) for v in range(n)) / float(c @ c))
    print(json.dumps({"n": n, "ndet": s.ndet, "blocks": s.nblocks, "ms (host wall, result copied back)": out,
                      "env": {k: v for k, v in os.environ.items() if k.startswith("F2_")}}), flush=True)


if __name__ == "__main__":
    main()
