#!/usr/bin/env python
"""alpha same-spin time per column for 8 alpha electrons in 16 orbitals and nb = 8 / 7 / 6 beta electrons: rows of
12870 (= 2 mod 4), 11440 (= 0 mod 4) and 8008 (= 0 mod 4) doubles -- does the 32-byte alignment of the rows matter?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import forte2_b200 as f2
from forte2_b200 import synthetic

E, H, V = synthetic.integrals(16, "S")
for nb in (8, 7, 6, 5):
    s = f2.CIStrings(8, nb, 0, [[0] * 16], [], [])
    b = f2.CISigmaBuilder(s, E, H, V, 0)
    b.set_memory(64 * 1024)
    c = torch.randn(s.ndet, dtype=torch.float64, device="cuda")
    sg = torch.empty_like(c)
    for _ in range(3):
        b.Hamiltonian(c, sg)
    torch.cuda.synchronize()
    ms = b.last_build_ms()
    print(f"nb={nb} nbs={s.nbs} (mod 4 = {s.nbs % 4}) alpha same-spin {ms[1]:.3f} ms = {ms[1] / s.nbs * 1e3:.4f} us per column; "
          f"opposite spin {ms[3]:.3f} ms = {ms[3] / s.nbs * 1e3:.4f} us per column", flush=True)
