"""Generate the golden fixtures from the reference's own C++ (oracle/_ref, built from /root/reference).

Run in the build container:  python tests/golden/make_golden.py
Writes tests/golden/lists_<space>.npz (string tables + every replacement list, reference order) and
tests/golden/sigma_<space>_<variant>.npz (integrals are regenerated from the seed; the file stores the
trial vector, sigma from Harrison-Zarrabian, determinant energies and, for tiny spaces, the dense
Slater-rules Hamiltonian)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import SMALL_SPACES  # noqa: E402
from forte2_b200 import synthetic  # noqa: E402
from oracle import ref  # noqa: E402

LIST_SPACES = ["cas66", "cas21_4", "sym2_11", "gas2_20", "gas3_44", "gas3_53", "gas343_54", "cas66_d2h",
               "gas2_sym4", "one_alpha", "empty"]
SIGMA_SPACES = ["cas66", "cas21_4", "sym2_11", "gas3_53", "gas343_54", "cas66_d2h", "cas66_cs1", "gas2_sym4",
                "cas75_odd", "one_each", "beta_only", "cas88"]


def main():
    for name in LIST_SPACES:
        s = ref.RefCIStrings(*SMALL_SPACES[name])
        out = {"info": np.array([s.na, s.nb, s.nas, s.nbs, s.ndet, s.norb, s.nirrep, s.ngas_spaces, s.nblocks,
                                 s.symmetry], dtype=np.int64),
               "blocks": s.blocks(), "occ_a": s.occupations(0), "occ_b": s.occupations(1),
               "occ_pairs": s.occupations(2)}
        for kind in ref.KINDS:
            for spin in (0, 1):
                out[f"{kind}_{spin}"] = s.export(kind, spin).reshape(-1)
        np.savez_compressed(os.path.join(HERE, f"lists_{name}.npz"), **out)
    for name in SIGMA_SPACES:
        space = SMALL_SPACES[name]
        n = sum(len(x) for x in space[3])
        s = ref.RefCIStrings(*space)
        for variant in ("S", "four_fold"):
            E, H, V = synthetic.four_fold_integrals(n) if variant == "four_fold" else synthetic.integrals(n, variant)
            b = ref.RefCISigmaBuilder(s, E, H, V)
            b.set_algorithm("hz")
            c = synthetic.trial_vector(s.ndet)
            sig = np.zeros_like(c)
            b.Hamiltonian(c, sig)
            out = {"c": c, "sigma_hz": sig, "hdiag": b.det_energies(), "dets": s.determinants()}
            if s.ndet <= 500:
                Hd = b.dense_hamiltonian()
                assert np.linalg.norm(Hd @ c - sig) <= 1e-13 * max(1.0, np.linalg.norm(sig))
                out["evals"] = np.linalg.eigvalsh(Hd)[:4]
            np.savez_compressed(os.path.join(HERE, f"sigma_{name}_{variant}.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
