"""Determinant-basis CI solver on the GPU: the caller of the sigma builder.

Mirrors the iterative path of forte2's `_CISingleStateSolver` (forte2/ci/ci.py:171-333, 411-452) in
its determinant-basis form (forte2/ci/rel_ci.py:96-114 — no CSF transform; at >= 1e8 determinants the
host-side CSF tables of the reference do not fit):

    strings -> CISigmaBuilder -> Hdiag (determinant energies, device) -> guess vectors from the
    lowest-energy determinants (host Slater rules on <= ndets_per_guess * guess_per_root * nroot
    determinants) -> DavidsonLiuSolver (device) with sigma_builder(B, S) looping Hamiltonian(B[i], S[i]).

Everything O(ndet) stays in HBM; the host sees the small guess Hamiltonian and the subspace problem.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .ci_sigma_builder import CISigmaBuilder
from .ci_strings import CIStrings
from .davidson import DavidsonLiuParams, DavidsonLiuSolver


def slater_matrix(strings: CIStrings, E, H, V, addresses) -> np.ndarray:
    """<I|H|J> for the determinants at `addresses` (host; forte2/determinant/slater_rules.cc:123-245)."""
    idx = np.ascontiguousarray(addresses, dtype=np.int64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    out = np.zeros((idx.size, idx.size))
    _lib.check(_lib.lib().f2_slater_matrix(strings._h, float(E), H.ctypes.data_as(_lib.f64p),
                                           V.ctypes.data_as(_lib.f64p), idx.ctypes.data_as(_lib.i64p),
                                           int(idx.size), out.ctypes.data_as(_lib.f64p)))
    return out


class CISolver:
    """One CI state (nroot lowest roots of one determinant space) on one B200.

    Parameters follow forte2/ci/ci.py:30-79 where they exist there: `nroot`, `davidson_liu_params`,
    `ci_algorithm` ("hz" | "kh" ...), `ci_builder_memory` (MB), `energy_shift`, `log_level`.
    """

    def __init__(self, strings: CIStrings, E, H, V, nroot=1, davidson_liu_params=None, ci_algorithm="hz",
                 ci_builder_memory=1024, energy_shift=None, log_level=3, device=None, log=None, builder=None):
        self.ci_strings = strings
        self.E0, self.H, self.V = float(E), np.asarray(H, dtype=np.float64), np.asarray(V, dtype=np.float64)
        self.nroot = int(nroot)
        self.davidson_liu_params = davidson_liu_params or DavidsonLiuParams()
        self.ci_algorithm, self.ci_builder_memory = ci_algorithm, int(ci_builder_memory)
        self.energy_shift, self.log_level, self.device, self._log = energy_shift, log_level, device, log
        self.ci_sigma_builder = builder  # an existing CISigmaBuilder for the same strings/integrals may be reused
        self.eigensolver = None
        self.executed = False

    # forte2/ci/ci.py:411-452
    def _build_guess_vectors(self, hdiag):
        import torch
        ndet = self.ci_strings.ndet
        p = self.davidson_liu_params
        self.num_guess_states = min(p.guess_per_root * self.nroot, ndet)
        nguess_dets = min(p.ndets_per_guess * self.num_guess_states, ndet)
        key = hdiag if self.energy_shift is None else (hdiag - self.energy_shift).abs()
        indices = torch.topk(key, nguess_dets, largest=False, sorted=True).indices
        idx = indices.cpu().numpy().astype(np.int64)
        Hguess = slater_matrix(self.ci_strings, self.E0, self.H, self.V, idx)
        _, evecs = np.linalg.eigh(Hguess)
        guess = torch.zeros((self.num_guess_states, ndet), dtype=torch.float64, device=hdiag.device)
        coeff = torch.from_numpy(np.ascontiguousarray(evecs[:, : self.num_guess_states].T)).to(hdiag.device)
        guess[:, indices] = coeff
        return guess

    def run(self):
        import torch
        s = self.ci_strings
        if s.ndet == 0:
            raise ValueError("No determinants could be generated for the given state and orbitals.")
        if self.ci_sigma_builder is None:
            self.ci_sigma_builder = CISigmaBuilder(s, self.E0, self.H, self.V, self.log_level, device=self.device)
            self.ci_sigma_builder.set_memory(self.ci_builder_memory)
        b = self.ci_sigma_builder
        b.set_algorithm(self.ci_algorithm.lower())
        hdiag = b.form_Hdiag_det()
        if s.ndet == 1:
            self.evals = hdiag[:1].cpu().numpy()
            self.evecs = torch.ones((1, 1), dtype=torch.float64, device=hdiag.device)
            self.E = self.evals
            self.executed = True
            return self
        if self.eigensolver is None:
            self.eigensolver = DavidsonLiuSolver(
                size=s.ndet, nroot=self.nroot, davidson_liu_params=self.davidson_liu_params,
                energy_shift=self.energy_shift, log_level=self.log_level, device=b.device, log=self._log)
            self.eigensolver.add_guesses(self._build_guess_vectors(hdiag))
        self.eigensolver.add_h_diag(hdiag)

        def sigma_builder(Bblock, Sblock):
            for i in range(Bblock.shape[0]):
                b.Hamiltonian(Bblock[i], Sblock[i])

        self.eigensolver.add_sigma_builder(sigma_builder)
        self.evals, self.evecs = self.eigensolver.solve()
        self.E = self.evals
        self.converged = self.eigensolver.converged
        self.executed = True
        return self
