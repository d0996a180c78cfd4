"""Quick device timings (not the bench): cuBLAS DGEMM peak and sigma build times."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import forte2_b200 as f2
from forte2_b200 import synthetic

def dgemm_peak():
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"cuBLAS DGEMM {n}^3: {2*n**3/best/1e9:.1f} TFLOP/s ({best:.2f} ms)", flush=True)

dgemm_peak()
for n in [int(x) for x in (sys.argv[1:] or ["12", "14"])]:
    s = f2.CIStrings(n // 2, n // 2, 0, [[0] * n], [], [])
    E, H, V = synthetic.integrals(n, "S")
    t = time.time()
    b = f2.CISigmaBuilder(s, E, H, V, 0)
    b.set_memory(64 * 1024)
    tb = time.time() - t
    c = torch.randn(s.ndet, dtype=torch.float64, device="cuda")
    c /= c.norm()
    sg = torch.empty_like(c)
    for _ in range(3):
        b.Hamiltonian(c, sg)
    torch.cuda.synchronize()
    ms = b.last_build_ms()
    print(f"CAS({n},{n}) ndet={s.ndet} builder {tb:.2f}s  last build ms [H0, aa, bb, ab, total] = "
          + ", ".join(f"{x:.3f}" for x in ms), flush=True)
