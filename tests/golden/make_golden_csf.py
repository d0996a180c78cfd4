"""Generate the CSF golden fixtures from the reference's own CISpinAdapter / form_Hdiag_csf
(oracle/_ref, built from /root/reference).

Run in the build container:  python tests/golden/make_golden_csf.py
Writes tests/golden/csf_<space>_S<twoS>.npz: the coupling table (bounds, det, coeff), nconf-independent
counts, the transforms of a seeded vector in both directions and the default CSF preconditioner on the
synthetic integrals (variant S)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import SMALL_SPACES  # noqa: E402
from forte2_b200 import synthetic  # noqa: E402
from oracle import ref  # noqa: E402

CSF_CASES = [("cas66", 0), ("cas66", 2), ("cas66_d2h", 0), ("gas343_54", 1), ("gas343_54", 3), ("cas75_odd", 1),
             ("cas88", 0), ("cas88", 4), ("gas2_sym4", 1), ("one_each", 0), ("full_shell", 0), ("gas3_53", 2)]


def main():
    for name, twoS in CSF_CASES:
        space = SMALL_SPACES[name]
        n = sum(len(x) for x in space[3])
        s = ref.RefCIStrings(*space)
        a = ref.RefCISpinAdapter(s, twoS, space[0] - space[1])
        bounds, det, coeff = a.table()
        E, H, V = synthetic.integrals(n, "S")
        b = ref.RefCISigmaBuilder(s, E, H, V)
        rng = np.random.default_rng(7)
        c_csf = rng.normal(size=a.ncsf)
        c_det = rng.normal(size=s.ndet)
        to_det = np.zeros(s.ndet)
        a.csf_C_to_det_C(c_csf, to_det)
        to_csf = np.zeros(a.ncsf)
        a.det_C_to_csf_C(c_det, to_csf)
        np.savez_compressed(os.path.join(HERE, f"csf_{name}_S{twoS}.npz"), bounds=bounds, det=det, coeff=coeff,
                            c_csf=c_csf, c_det=c_det, to_det=to_det, to_csf=to_csf, hdiag_csf=a.hdiag_csf(b, False))
    print("CSF golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
