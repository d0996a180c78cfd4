"""Generate tests/golden/roots_<space>_<variant>.npz: CI roots from the reference's own C++ sigma
builder (oracle/_ref) driven by the reference's own, unmodified DavidsonLiuSolver
(/root/reference/forte2/helpers/davidsonliu.py loaded by oracle/refdl.py) in the determinant basis,
with the guess of forte2/ci/ci.py:411-452 (lowest determinant energies, dense Slater rules).
For spaces up to 5000 determinants the exact spectrum of the reference's dense Hamiltonian is stored too.

Run in the build container:  python tests/golden/make_golden_roots.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from cases import SMALL_SPACES  # noqa: E402
from forte2_b200 import synthetic  # noqa: E402
from oracle import ref, refdl  # noqa: E402

CASES = [("cas66", "D"), ("cas66", "S"), ("cas88", "D"), ("gas343_55", "D"), ("cas66_d2h", "D"), ("cas75_odd", "D"),
         ("cas1010", "D")]
SPACES = dict(SMALL_SPACES, cas1010=(5, 5, 0, [[0] * 10], [], []))


def reference_roots(space, ints, nroot):
    Solver, Params = refdl.load()
    E, H, V = ints
    s = ref.RefCIStrings(*space)
    b = ref.RefCISigmaBuilder(s, E, H, V)
    b.set_algorithm("hz")
    b.set_memory(4096)
    hdiag = b.det_energies()
    p = Params()
    nstates = min(p.guess_per_root * nroot, s.ndet)
    ndets = min(p.ndets_per_guess * nstates, s.ndet)
    idx = np.argsort(hdiag, kind="stable")[:ndets]
    # guess Hamiltonian: columns of H through sigma on unit vectors (the reference builder itself)
    Hg = np.zeros((ndets, ndets))
    for j, J in enumerate(idx):
        e = np.zeros(s.ndet)
        e[J] = 1.0
        sg = np.zeros(s.ndet)
        b.Hamiltonian(e, sg)
        Hg[:, j] = sg[idx]
    _, ev = np.linalg.eigh(0.5 * (Hg + Hg.T))
    guess = np.zeros((s.ndet, nstates))
    guess[idx, :] = ev[:, :nstates]
    solver = Solver(size=s.ndet, nroot=nroot, davidson_liu_params=p, log_level=5)
    solver.add_h_diag(hdiag)
    solver.add_guesses(guess)
    bd, sd = np.zeros(s.ndet), np.zeros(s.ndet)

    def sigma_builder(B, S):
        for i in range(B.shape[1]):
            bd[:] = B[:, i]
            b.Hamiltonian(bd, sd)
            S[:, i] = sd

    solver.add_sigma_builder(sigma_builder)
    evals, _ = solver.solve()
    assert solver.converged
    exact = None
    if s.ndet <= 5000:
        exact = np.linalg.eigvalsh(b.dense_hamiltonian())[:6]
    return np.array(evals), solver.iter + 1, exact


def main():
    for name, variant in CASES:
        space = SPACES[name]
        n = sum(len(x) for x in space[3])
        ints = synthetic.integrals(n, variant)
        out = {}
        for nroot in (1, 3):
            ev, iters, exact = reference_roots(space, ints, nroot)
            out[f"evals_nroot{nroot}"] = ev
            out[f"iters_nroot{nroot}"] = np.array(iters)
            if exact is not None:
                out["exact"] = exact
                assert np.max(np.abs(ev - exact[:nroot])) < 1e-9, (name, variant, ev, exact)
        np.savez_compressed(os.path.join(HERE, f"roots_{name}_{variant}.npz"), **out)
        print(name, variant, {k: v for k, v in out.items()})


if __name__ == "__main__":
    main()
