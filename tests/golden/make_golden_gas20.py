"""Golden fixture for the 1-GPU validation twin of BASELINE.json configs[4] (3 GAS spaces, 20 orbitals;
SURVEY.md section 8d): na = nb = 10, GAS sizes 6/8/6, gas_min = [10,0,0], gas_max = [12,16,2]:
22 065 484 determinants in 36 blocks.  sigma from the reference's Knowles-Handy builder (oracle/_ref; valid
with one irrep) on the synthetic integrals (variant S) and the seeded trial vector.  The full
vector is 176 MB, so the fixture keeps 200 000 seeded sample positions, the norm and <c|sigma>.
Takes about an hour on 8 cores:  python tests/golden/make_golden_gas20.py [sigma.npy]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from forte2_b200 import synthetic  # noqa: E402

SPACE = (10, 10, 0, [[0] * 6, [0] * 8, [0] * 6], [10, 0, 0], [12, 16, 2])


def main():
    if len(sys.argv) > 1:
        sg = np.load(sys.argv[1])
    else:
        from oracle import ref
        rs = ref.RefCIStrings(*SPACE)
        E, H, V = synthetic.integrals(20, "S")
        rb = ref.RefCISigmaBuilder(rs, E, H, V)
        rb.set_algorithm(os.environ.get("F2_GOLDEN_ALGORITHM", "kh"))  # "hz" did not finish in 90 minutes on 8 cores either
        rb.set_memory(16384)
        c = synthetic.trial_vector(rs.ndet)
        sg = np.zeros_like(c)
        rb.Hamiltonian(c, sg)
    c = synthetic.trial_vector(len(sg))
    idx = np.sort(np.random.default_rng(20).choice(len(sg), size=200000, replace=False))
    np.savez_compressed(os.path.join(HERE, "gas20twin_S.npz"), ndet=np.array(len(sg)), idx=idx, sigma_at_idx=sg[idx],
                        norm=np.array(np.linalg.norm(sg)), c_dot_sigma=np.array(float(c @ sg)))
    print("written", len(sg), np.linalg.norm(sg))


if __name__ == "__main__":
    main()
