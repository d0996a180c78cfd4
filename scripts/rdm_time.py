#!/usr/bin/env python
"""Device time of every RDM spin block of CAS(n,n) through f2_sigma_rdm (CUDA events around the C-ABI call, result left
on the device): usage rdm_time.py n [memory_mb] [max body]; RDM_KINDS=ab_2rdm,... restricts the blocks.  Prints one JSON line per block: ms, executed DMMA flops
(2 M_pad^2 K with the 128-row tile padding), useful flops (2 M^2 K), TFLOP/s of both."""
import json
import os
import sys
from math import comb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import forte2_b200 as f2
    from forte2_b200 import synthetic, _lib
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    mem = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    body = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    E, H, V = synthetic.integrals(n, "S")
    na = nb = n // 2
    strings = f2.CIStrings(na, nb, 0, [[0] * n], [], [])
    b = f2.CISigmaBuilder(strings, E, H, V, 0)
    b.set_memory(mem)
    c = torch.from_numpy(synthetic.trial_vector(strings.ndet)).cuda()
    L = _lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    kinds = {"a_1rdm": (0, 1, 0), "b_1rdm": (1, 0, 1), "aa_2rdm": (3, 2, 0), "bb_2rdm": (4, 0, 2), "ab_2rdm": (5, 1, 1),
             "aaa_3rdm": (8, 3, 0), "aab_3rdm": (9, 2, 1), "abb_3rdm": (10, 1, 2), "bbb_3rdm": (11, 0, 3)}
    for name, (kind, la, lb) in kinds.items():
        if la + lb > body or (os.environ.get("RDM_KINDS") and name not in os.environ["RDM_KINDS"].split(",")):
            continue
        M = comb(n, la) * comb(n, lb)
        K = comb(n, na - la) * comb(n, nb - lb)
        out = torch.empty(M * M, dtype=torch.float64, device="cuda")
        _lib.check(L.f2_sigma_rdm(b._h, kind, c.data_ptr(), c.data_ptr(), out.data_ptr(), stream))  # warm-up, uploads
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.f2_sigma_rdm(b._h, kind, c.data_ptr(), c.data_ptr(), out.data_ptr(), stream))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        Mp = (M + 127) // 128 * 128
        tr = float(out.view(M, M).diagonal().sum())
        cc = sum(float(c[i:i + (1 << 30)] @ c[i:i + (1 << 30)]) for i in range(0, c.numel(), 1 << 30))  # dot stops at 2^31 - 1
        want_tr = comb(na, la) * comb(nb, lb) * cc
        print(json.dumps({"n": n, "block": name, "M": M, "K": K, "ms": round(ms, 3), "launches": b.launch_count(),
                          "useful_tflops": round(2 * M * M * K / ms / 1e9, 2),
                          "executed_tflops": round(2 * Mp * Mp * K / ms / 1e9, 2),
                          "D_bytes": 8 * M * K, "trace_err": abs(tr - want_tr) / want_tr}), flush=True)


if __name__ == "__main__":
    main()
