"""Generate tests/golden/rootscsf_<space>_S<twoS>.npz: spin-adapted CI roots from the reference's own
pieces wired as forte2/ci/ci.py:147-235 wires them: CISpinAdapter (oracle/_ref), form_Hdiag_csf
preconditioner, the guess of ci.py:411-452 in the CSF basis, the closure csf_C_to_det_C ->
CISigmaBuilder::Hamiltonian -> det_C_to_csf_C, and the reference's unmodified DavidsonLiuSolver
(oracle/refdl.py).  The exact lowest eigenvalues of the CSF-projected dense Hamiltonian are stored too.

Run in the build container:  python tests/golden/make_golden_roots_csf.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import SMALL_SPACES  # noqa: E402
from forte2_b200 import synthetic  # noqa: E402
from oracle import ref, refdl  # noqa: E402

CASES = [("cas66", 0), ("cas66", 2), ("cas75_odd", 1), ("cas66_d2h", 0), ("gas343_54", 1), ("cas88", 0)]


def reference_roots_csf(space, twoS, ints, nroot):
    Solver, Params = refdl.load()
    E, H, V = ints
    s = ref.RefCIStrings(*space)
    b = ref.RefCISigmaBuilder(s, E, H, V)
    b.set_algorithm("hz")
    a = ref.RefCISpinAdapter(s, twoS, space[0] - space[1])
    bounds, det, coeff = a.table()
    T = np.zeros((a.ncsf, s.ndet))
    for i in range(a.ncsf):
        T[i, det[bounds[i]:bounds[i + 1]]] = coeff[bounds[i]:bounds[i + 1]]
    Hcsf = T @ b.dense_hamiltonian() @ T.T
    exact = np.linalg.eigvalsh(0.5 * (Hcsf + Hcsf.T))[:6]
    hdiag = a.hdiag_csf(b, False)
    p = Params()
    nstates = min(p.guess_per_root * nroot, a.ncsf)
    nguess = min(p.ndets_per_guess * nstates, a.ncsf)
    idx = np.argsort(hdiag, kind="stable")[:nguess]
    Hg = Hcsf[np.ix_(idx, idx)]
    _, ev = np.linalg.eigh(0.5 * (Hg + Hg.T))
    guess = np.zeros((a.ncsf, nstates))
    guess[idx, :] = ev[:, :nstates]
    solver = Solver(size=a.ncsf, nroot=nroot, davidson_liu_params=p, log_level=5)
    solver.add_h_diag(hdiag)
    solver.add_guesses(guess)
    b_det, s_det = np.zeros(s.ndet), np.zeros(s.ndet)
    b_csf, s_csf = np.zeros(a.ncsf), np.zeros(a.ncsf)

    def sigma_builder(B, S):
        for i in range(B.shape[1]):
            b_csf[:] = B[:, i]
            a.csf_C_to_det_C(b_csf, b_det)
            b.Hamiltonian(b_det, s_det)
            a.det_C_to_csf_C(s_det, s_csf)
            S[:, i] = s_csf

    solver.add_sigma_builder(sigma_builder)
    evals, _ = solver.solve()
    assert solver.converged
    return np.array(evals), exact, a.ncsf


def main():
    for name, twoS in CASES:
        space = SMALL_SPACES[name]
        n = sum(len(x) for x in space[3])
        ints = synthetic.integrals(n, "D")
        out = {}
        for nroot in (1, 3):
            ev, exact, ncsf = reference_roots_csf(space, twoS, ints, nroot)
            assert np.max(np.abs(ev - exact[:nroot])) < 1e-9, (name, twoS, ev, exact)
            out[f"evals_nroot{nroot}"] = ev
            out["exact"] = exact
            out["ncsf"] = np.array(ncsf)
        np.savez_compressed(os.path.join(HERE, f"rootscsf_{name}_S{twoS}.npz"), **out)
        print(name, twoS, out["evals_nroot3"], out["exact"][:3])


if __name__ == "__main__":
    main()
