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


def spin2_matrix(strings: CIStrings, addresses) -> np.ndarray:
    """<I|S^2|J> for the determinants at `addresses` (host; det.spin2,
    forte2/determinant/determinant_helpers.cc:154-203)."""
    idx = np.ascontiguousarray(addresses, dtype=np.int64)
    out = np.zeros((idx.size, idx.size))
    _lib.check(_lib.lib().f2_spin2_matrix(strings._h, idx.ctypes.data_as(_lib.i64p), int(idx.size),
                                          out.ctypes.data_as(_lib.f64p)))
    return out


def spin_complete(strings: CIStrings, addresses) -> np.ndarray:
    """The given determinants plus every determinant with the same doubly / singly occupied orbitals and the same
    Ms (forte2/sci/sci.py:543-565, _generate_spin_complement_pairs): S^2 is block diagonal over such sets.  Keeps
    the given order, complements appended in first-seen order."""
    from itertools import combinations
    out, seen = [], set()
    for adr in np.asarray(addresses, dtype=np.int64):
        a, b = strings.determinant(int(adr))
        docc, socc = a & b, a ^ b
        open_orbs = [p for p in range(strings.norb) if (socc >> p) & 1]
        nao = bin(a & socc).count("1")
        first = [int(adr)]
        for alpha in combinations(open_orbs, nao):
            wa = docc
            for p in alpha:
                wa |= 1 << p
            wb = docc | (socc & ~wa)
            if (wa, wb) != (a, b):
                first.append(strings.determinant_index((wa, wb)))
        for x in first:
            if x not in seen:
                seen.add(x)
                out.append(x)
    return np.array(out, dtype=np.int64)


def spin_adapted_guess(strings: CIStrings, E, H, V, addresses, multiplicity, nstates):
    """Spin-pure guess vectors in the space of the given determinants, as forte2's determinant-basis driver makes them
    (forte2/sci/sci.py:405-455): the determinant set is made spin complete, S^2 is diagonalised in it, the guess
    Hamiltonian is diagonalised inside the eigenspace of S(S+1), and the other S^2 eigenvectors are returned to be
    projected out of every trial vector.  Returns (addresses, guess (n, nstates'), wrong_spin (n, nwrong))."""
    idx = spin_complete(strings, addresses)
    S2 = spin2_matrix(strings, idx)
    svals, svecs = np.linalg.eigh(S2)
    S = 0.5 * (multiplicity - 1)
    close = np.isclose(svals, S * (S + 1.0), atol=1e-8)
    if not close.any():
        raise ValueError(f"no guess determinant combination with S(S+1) = {S * (S + 1.0)}: the multiplicity "
                         f"{multiplicity} is not compatible with na - nb = {strings.na - strings.nb}")
    sub = svecs[:, close]
    Hg = slater_matrix(strings, E, H, V, idx)
    Hs = sub.T @ Hg @ sub
    _, ev = np.linalg.eigh(0.5 * (Hs + Hs.T))
    return idx, sub @ ev[:, :min(nstates, sub.shape[1])], svecs[:, ~close]


class CISolver:
    """One CI state (nroot lowest roots of one determinant space) on one B200.

    Parameters follow forte2/ci/ci.py:30-79 where they exist there: `nroot`, `davidson_liu_params`,
    `ci_algorithm` ("hz" | "kh" | "exact" ...), `ci_builder_memory` (MB), `energy_shift`, `log_level`.

    With `spin_adapter` (a prepared forte2_b200.CISpinAdapter) the solve runs in the CSF basis exactly as
    forte2's default path does (forte2/ci/ci.py:147-165, 214-235): the Davidson vectors have ncsf elements,
    the preconditioner is form_Hdiag_csf, and every sigma build is csf_C_to_det_C -> Hamiltonian ->
    det_C_to_csf_C, all on device vectors.  Without it the solve runs in the determinant basis
    (forte2/ci/rel_ci.py:96-114), which is the only option once the coupling table no longer fits.

    Determinant-basis spin control (`multiplicity` given, no spin adapter): what keeps the roots of a large solve
    in the wanted spin sector.  As forte2's own determinant-basis driver does (forte2/sci/sci.py:405-455,543-565)
    the guess determinants are made spin complete, the guesses are eigenvectors of S^2 with eigenvalue S(S+1) and
    the other S^2 eigenvectors of the guess space are projected out of every sigma and correction vector; in addition
    the Davidson preconditioner is the configuration-averaged diagonal (f2_sigma_hdiag_spin_averaged), which commutes
    with S^2, so the corrections do not leave the spin sector.
    """

    def __init__(self, strings: CIStrings, E, H, V, nroot=1, davidson_liu_params=None, ci_algorithm="hz",
                 ci_builder_memory=1024, energy_shift=None, log_level=3, device=None, log=None, builder=None,
                 spin_adapter=None, multiplicity=None):
        self.spin_adapter = spin_adapter
        self.multiplicity = None if multiplicity is None else int(multiplicity)
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
        if self.multiplicity is not None:
            idx, coeff, wrong = spin_adapted_guess(self.ci_strings, self.E0, self.H, self.V, idx, self.multiplicity,
                                                   self.num_guess_states)
            if coeff.shape[1] < min(self.nroot, self.num_guess_states):
                raise ValueError(f"only {coeff.shape[1]} guess vectors of multiplicity {self.multiplicity} in the "
                                 f"space of the {idx.size} guess determinants; increase ndets_per_guess")
            self.num_guess_states = coeff.shape[1]
            self._guess_addresses, self._wrong_spin = idx, wrong
            indices = torch.from_numpy(idx).to(hdiag.device)
            guess = torch.zeros((self.num_guess_states, ndet), dtype=torch.float64, device=hdiag.device)
            guess[:, indices] = torch.from_numpy(np.ascontiguousarray(coeff.T)).to(hdiag.device)
            return guess
        Hguess = slater_matrix(self.ci_strings, self.E0, self.H, self.V, idx)
        _, evecs = np.linalg.eigh(Hguess)
        guess = torch.zeros((self.num_guess_states, ndet), dtype=torch.float64, device=hdiag.device)
        coeff = torch.from_numpy(np.ascontiguousarray(evecs[:, : self.num_guess_states].T)).to(hdiag.device)
        guess[:, indices] = coeff
        return guess

    # forte2/ci/ci.py:405-452 with _basis_matrix_element = slater_rules_csf
    def _build_guess_vectors_csf(self, hdiag):
        import torch
        sa = self.spin_adapter
        p = self.davidson_liu_params
        self.num_guess_states = min(p.guess_per_root * self.nroot, sa.ncsf)
        nguess = min(p.ndets_per_guess * self.num_guess_states, sa.ncsf)
        key = hdiag if self.energy_shift is None else (hdiag - self.energy_shift).abs()
        indices = torch.topk(key, nguess, largest=False, sorted=True).indices
        rows = [sa.csf(int(i)) for i in indices.cpu().numpy()]
        dets = sorted({d for row in rows for d, _ in row})
        pos = {d: k for k, d in enumerate(dets)}
        C = np.zeros((len(dets), nguess))
        for j, row in enumerate(rows):
            for d, c in row:
                C[pos[d], j] += c
        Hdet = slater_matrix(self.ci_strings, self.E0, self.H, self.V, np.array(dets, dtype=np.int64))
        Hguess = C.T @ Hdet @ C
        _, evecs = np.linalg.eigh(0.5 * (Hguess + Hguess.T))
        guess = torch.zeros((self.num_guess_states, sa.ncsf), dtype=torch.float64, device=hdiag.device)
        coeff = torch.from_numpy(np.ascontiguousarray(evecs[:, : self.num_guess_states].T)).to(hdiag.device)
        guess[:, indices] = coeff
        return guess

    def evecs_det(self):
        """the converged roots in the determinant basis, (nroot, ndet) CUDA tensor"""
        import torch
        if self.spin_adapter is None:
            return self.evecs
        out = torch.empty((self.evecs.shape[0], self.ci_strings.ndet), dtype=torch.float64, device=self.evecs.device)
        for i in range(self.evecs.shape[0]):
            self.spin_adapter.csf_C_to_det_C(self.evecs[i].contiguous(), out[i])
        return out

    # ---- reduced density matrices of the converged roots (forte2/ci/ci.py:476-760: make_1rdm ... make_sf_3cumulant);
    # the roots stay on the device, every spin block comes from the builder's kernels
    def _root_pair(self, left_root: int, right_root):
        if not self.executed:
            raise RuntimeError("CI solver has not been executed yet; call run() first")
        ev = self.evecs_det()
        left = ev[int(left_root)].contiguous()
        return left, (left if right_root is None else ev[int(right_root)].contiguous())

    def make_sd_1rdm(self, left_root: int, right_root=None):
        """spin-dependent 1-RDMs (a, b)"""
        l, r = self._root_pair(left_root, right_root)
        return self.ci_sigma_builder.a_1rdm(l, r), self.ci_sigma_builder.b_1rdm(l, r)

    def make_sd_2rdm(self, left_root: int, right_root=None):
        """spin-dependent 2-RDMs (aa, ab, bb)"""
        l, r = self._root_pair(left_root, right_root)
        b = self.ci_sigma_builder
        return b.aa_2rdm(l, r), b.ab_2rdm(l, r), b.bb_2rdm(l, r)

    def make_sd_3rdm(self, left_root: int, right_root=None):
        """spin-dependent 3-RDMs (aaa, aab, abb, bbb)"""
        l, r = self._root_pair(left_root, right_root)
        b = self.ci_sigma_builder
        return b.aaa_3rdm(l, r), b.aab_3rdm(l, r), b.abb_3rdm(l, r), b.bbb_3rdm(l, r)

    def make_sf_1rdm(self, left_root: int, right_root=None):
        return self.ci_sigma_builder.sf_1rdm(*self._root_pair(left_root, right_root))

    def make_sf_2rdm(self, left_root: int, right_root=None):
        return self.ci_sigma_builder.sf_2rdm(*self._root_pair(left_root, right_root))

    def make_sf_3rdm(self, left_root: int, right_root=None):
        return self.ci_sigma_builder.sf_3rdm(*self._root_pair(left_root, right_root))

    def make_sf_2cumulant(self, left_root: int, right_root=None):
        return self.ci_sigma_builder.sf_2cumulant(*self._root_pair(left_root, right_root))

    def make_sf_3cumulant(self, left_root: int, right_root=None):
        return self.ci_sigma_builder.sf_3cumulant(*self._root_pair(left_root, right_root))

    # the non-relativistic solver's make_1rdm ... make_3cumulant are the spin-free ones (forte2/ci/ci.py:476-574)
    make_1rdm, make_2rdm, make_3rdm = make_sf_1rdm, make_sf_2rdm, make_sf_3rdm
    make_2cumulant, make_3cumulant = make_sf_2cumulant, make_sf_3cumulant

    def _run_csf(self, b):
        import torch
        s, sa = self.ci_strings, self.spin_adapter
        if sa.ndet != s.ndet:
            raise ValueError("the spin adapter was prepared for a different determinant space")
        if sa.ncsf == 0:
            raise ValueError("No CSFs could be generated for the given state and orbitals.")
        hdiag = b.form_Hdiag_csf(None, sa)
        if sa.ncsf == 1:
            self.evals = self.E = hdiag[:1].cpu().numpy()
            self.evecs = torch.ones((1, 1), dtype=torch.float64, device=hdiag.device)
            self.converged = self.executed = True
            return self
        if self.eigensolver is None:
            self.eigensolver = DavidsonLiuSolver(
                size=sa.ncsf, nroot=self.nroot, davidson_liu_params=self.davidson_liu_params,
                energy_shift=self.energy_shift, log_level=self.log_level, device=b.device, log=self._log)
            self.eigensolver.add_guesses(self._build_guess_vectors_csf(hdiag))
        self.eigensolver.add_h_diag(hdiag)
        b_det = torch.empty(s.ndet, dtype=torch.float64, device=hdiag.device)
        sigma_det = torch.empty_like(b_det)

        def sigma_builder(Bblock, Sblock):  # forte2/ci/ci.py:228-233
            for i in range(Bblock.shape[0]):
                sa.csf_C_to_det_C(Bblock[i], b_det)
                b.Hamiltonian(b_det, sigma_det)
                sa.det_C_to_csf_C(sigma_det, Sblock[i])

        self.eigensolver.add_sigma_builder(sigma_builder)
        self.evals, self.evecs = self.eigensolver.solve()
        self.E = self.evals
        self.converged = self.eigensolver.converged
        self.executed = True
        return self

    def _do_exact_diagonalization(self, b):
        """ci_algorithm = "exact" (forte2/ci/ci.py:193-195,305-322; the reference's cross-check of its sigma
        algorithms, tests/ci/test_ci_exact_diagonalization.py): the dense Hamiltonian of the variational basis --
        CSFs with a spin adapter (CISigmaBuilder::form_H_csf), determinants without one (rel_ci.py:125-131) -- from
        host Slater rules, numpy eigh, energy_shift ordering.  Small spaces only: the builder's memory limit
        applies as in the reference.  Roots come back as rows of a CUDA tensor like those of the iterative solve."""
        import torch
        if self.spin_adapter is not None:
            Hd = b.form_H_csf(None, self.spin_adapter)
        else:
            n = self.ci_strings.ndet
            if n * n * 8 > self.ci_builder_memory * 1024 * 1024:
                raise RuntimeError("Not enough memory to form the full Hamiltonian matrix in the determinant "
                                   "basis. Increase the memory limit or use a different algorithm.")
            Hd = slater_matrix(self.ci_strings, self.E0, self.H, self.V, np.arange(n, dtype=np.int64))
            Hd = 0.5 * (Hd + Hd.T)
        evals, evecs = np.linalg.eigh(Hd)
        if self.energy_shift is not None:
            order = np.argsort(np.abs(evals - self.energy_shift))
            evals, evecs = evals[order], evecs[:, order]
        self.evals_full, self.evecs_full = evals, evecs
        k = min(self.nroot, evals.size)
        self.evals = self.E = evals[:k].copy()
        self.evecs = torch.from_numpy(np.ascontiguousarray(evecs[:, :k].T)).to(f"cuda:{b.device}")
        self.converged = self.executed = True
        return self

    def run(self):
        import torch
        s = self.ci_strings
        if s.ndet == 0:
            raise ValueError("No determinants could be generated for the given state and orbitals.")
        if self.ci_sigma_builder is None:
            self.ci_sigma_builder = CISigmaBuilder(s, self.E0, self.H, self.V, self.log_level, device=self.device)
            self.ci_sigma_builder.set_memory(self.ci_builder_memory)
        b = self.ci_sigma_builder
        if self.ci_algorithm.lower() == "exact":
            return self._do_exact_diagonalization(b)
        b.set_algorithm(self.ci_algorithm.lower())
        if self.spin_adapter is not None:
            return self._run_csf(b)
        hdiag = b.form_Hdiag_det(spin_averaged=self.multiplicity is not None)
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
            if self.multiplicity is not None and self._wrong_spin.shape[1] > 0:
                self.eigensolver.add_project_out_sparse(self._guess_addresses, self._wrong_spin.T)
        self.eigensolver.add_h_diag(hdiag)

        def sigma_builder(Bblock, Sblock):
            b.Hamiltonian(Bblock, Sblock)  # every new trial vector in one call (f2_sigma_apply_batch)

        self.eigensolver.add_sigma_builder(sigma_builder)
        self.evals, self.evecs = self.eigensolver.solve()
        self.E = self.evals
        self.converged = self.eigensolver.converged
        self.executed = True
        return self


class ShardedCISolver:
    """The same solve with the CI vector sharded by alpha-string rows over the ranks of a
    torch.distributed process group (one process per GPU; forte2_b200/dist.py).  No reference
    counterpart (forte2 is single-process); the roots equal the single-GPU roots.

    Every O(ndet) object (Hdiag, guesses, the Davidson subspace, sigma) exists only as the rank's row
    slab; the ranks agree on the guess determinants through an all-gather of their local candidates.
    """

    def __init__(self, strings: CIStrings, E, H, V, nroot=1, davidson_liu_params=None, ci_algorithm="hz",
                 ci_builder_memory=1024, log_level=3, device=None, log=None, builder=None, group=None,
                 form="auto", multiplicity=None, _sharded=None, _hdiag=None, _kernels=None):
        # _sharded / _hdiag / _kernels: test hooks (tests/ run the host logic over gloo with CPU stand-ins
        # for the GPU work); the product path below has no CPU fallback
        import torch.distributed as dist
        self.ci_strings = strings
        self.E0, self.H, self.V = float(E), np.asarray(H, dtype=np.float64), np.asarray(V, dtype=np.float64)
        self.nroot = int(nroot)
        self.davidson_liu_params = davidson_liu_params or DavidsonLiuParams()
        self.ci_algorithm, self.ci_builder_memory = ci_algorithm, int(ci_builder_memory)
        self.log_level, self.device, self._log = log_level, device, log
        self.ci_sigma_builder = builder
        # the Davidson kernels treat comm=None as "not sharded": always hand them a real group
        self.group, self.form = (group if group is not None else dist.group.WORLD), form
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.eigensolver = None
        self.sharded = _sharded
        self.multiplicity = None if multiplicity is None else int(multiplicity)  # spin control as in CISolver
        self._wrong_spin = None
        self._hdiag, self._kernels = _hdiag, _kernels
        self.executed = False

    def _build_guess_vectors(self, hdiag_loc):
        """forte2/ci/ci.py:411-452 on sharded vectors: the nguess lowest-energy determinants are the
        lowest of the union of every rank's nguess lowest."""
        import torch
        import torch.distributed as dist
        lay = self.sharded.layout
        ndet = self.ci_strings.ndet
        p = self.davidson_liu_params
        self.num_guess_states = min(p.guess_per_root * self.nroot, ndet)
        nguess_dets = min(p.ndets_per_guess * self.num_guess_states, ndet)
        k = min(nguess_dets, hdiag_loc.numel())
        vals = torch.full((nguess_dets,), float("inf"), dtype=torch.float64, device=hdiag_loc.device)
        addr = torch.full((nguess_dets,), -1, dtype=torch.int64, device=hdiag_loc.device)
        if k > 0:
            top = torch.topk(hdiag_loc, k, largest=False, sorted=True)
            vals[:k] = top.values
            addr[:k] = torch.from_numpy(lay.global_of_local(top.indices.cpu().numpy())).to(addr.device)
        all_vals = [torch.empty_like(vals) for _ in range(self.world)]
        all_addr = [torch.empty_like(addr) for _ in range(self.world)]
        dist.all_gather(all_vals, vals, group=self.group)
        dist.all_gather(all_addr, addr, group=self.group)
        v = torch.cat(all_vals).cpu().numpy()
        a = torch.cat(all_addr).cpu().numpy()
        keep = a >= 0
        v, a = v[keep], a[keep]
        order = np.lexsort((a, v))[:nguess_dets]  # by energy, ties by address: identical on every rank
        idx = a[order].astype(np.int64)
        if self.multiplicity is not None:
            idx, evecs, wrong = spin_adapted_guess(self.ci_strings, self.E0, self.H, self.V, idx, self.multiplicity,
                                                   self.num_guess_states)
            if evecs.shape[1] < min(self.nroot, self.num_guess_states):
                raise ValueError(f"only {evecs.shape[1]} guess vectors of multiplicity {self.multiplicity} in the "
                                 f"space of the {idx.size} guess determinants; increase ndets_per_guess")
            self.num_guess_states = evecs.shape[1]
            if wrong.shape[1] > 0:  # globally orthonormal (eigenvectors of S^2); every rank keeps its components
                loc = lay.local_of_global(idx)
                own = np.nonzero(loc >= 0)[0]
                self._wrong_spin = (loc[own], np.ascontiguousarray(wrong[own].T))
        else:
            Hguess = slater_matrix(self.ci_strings, self.E0, self.H, self.V, idx)
            _, evecs = np.linalg.eigh(Hguess)
        guess = torch.zeros((self.num_guess_states, lay.local_size), dtype=torch.float64, device=hdiag_loc.device)
        local = lay.local_of_global(idx)
        mine = np.nonzero(local >= 0)[0]
        if mine.size:
            coeff = np.ascontiguousarray(evecs[mine, : self.num_guess_states].T)
            guess[:, torch.from_numpy(local[mine]).to(guess.device)] = torch.from_numpy(coeff).to(guess.device)
        return guess

    def run(self):
        from .dist import ShardedSigma
        s = self.ci_strings
        if s.ndet == 0:
            raise ValueError("No determinants could be generated for the given state and orbitals.")
        b = self.ci_sigma_builder
        if self.sharded is None or self._hdiag is None:
            if b is None:
                b = self.ci_sigma_builder = CISigmaBuilder(s, self.E0, self.H, self.V, self.log_level,
                                                           device=self.device)
                b.set_memory(self.ci_builder_memory)
            b.set_algorithm(self.ci_algorithm.lower())
        if self.sharded is None:
            self.sharded = ShardedSigma(b, self.rank, self.world, group=self.group, form=self.form)
        sh = self.sharded
        hdiag = self._hdiag if self._hdiag is not None else b.form_Hdiag_det_part(
            self.rank, self.world, sh.local_size, spin_averaged=self.multiplicity is not None)
        if self.eigensolver is None:
            self.eigensolver = DavidsonLiuSolver(
                size=sh.local_size, nroot=self.nroot, davidson_liu_params=self.davidson_liu_params,
                log_level=self.log_level, device=hdiag.device, comm=self.group, global_size=s.ndet, log=self._log,
                _kernels=self._kernels)
            self.eigensolver.add_guesses(self._build_guess_vectors(hdiag))
            if self._wrong_spin is not None:
                self.eigensolver.add_project_out_sparse(*self._wrong_spin)
        self.eigensolver.add_h_diag(hdiag)

        def sigma_builder(Bblock, Sblock):
            for i in range(Bblock.shape[0]):
                sh.Hamiltonian(Bblock[i], Sblock[i])

        self.eigensolver.add_sigma_builder(sigma_builder)
        self.evals, self.evecs = self.eigensolver.solve()
        self.E = self.evals
        self.converged = self.eigensolver.converged
        self.executed = True
        return self


def _sharded_root_rdms(self, left_root: int, right_root=None):
    """spin-free 1- and 2-RDM of a root (pair) of a sharded solve: full copies of the vectors on every rank, the work of
    every spin block split over the ranks, one all-reduce per block (forte2_b200/dist.py sharded_rdms)"""
    from .ci_sigma_builder import assemble_sf_2rdm
    from .dist import sharded_rdms
    if not self.executed:
        raise RuntimeError("CI solver has not been executed yet; call run() first")
    left = self.evecs[int(left_root)].contiguous()
    right = None if right_root is None else self.evecs[int(right_root)].contiguous()
    blocks = sharded_rdms(self.sharded, self.ci_sigma_builder, left, right)
    n = self.ci_strings.norb
    g1 = blocks["a_1rdm"] + blocks["b_1rdm"]
    g2 = assemble_sf_2rdm(n, blocks["ab_2rdm"], blocks["aa_2rdm"] if n >= 2 else None, blocks["bb_2rdm"] if n >= 2 else None)
    return g1, g2


def _sharded_make_sf_1rdm(self, left_root: int, right_root=None):
    return _sharded_root_rdms(self, left_root, right_root)[0]


def _sharded_make_sf_2rdm(self, left_root: int, right_root=None):
    return _sharded_root_rdms(self, left_root, right_root)[1]


ShardedCISolver.make_sf_rdms = _sharded_root_rdms
ShardedCISolver.make_sf_1rdm = ShardedCISolver.make_1rdm = _sharded_make_sf_1rdm
ShardedCISolver.make_sf_2rdm = ShardedCISolver.make_2rdm = _sharded_make_sf_2rdm


def make_average_rdms(sub_solvers, weights, max_body: int = 2):
    """State-averaged spin-free RDMs over the roots of one or several solved CISolver objects
    (forte2/base_classes/ci_base.py:104-149: make_average_1rdm / make_average_2rdm / make_average_3rdm):
    weights[i][j] is the weight of root j of sub_solvers[i].  Returns (G1, G2) or, with max_body = 3, (G1, G2, G3)."""
    out = None
    for solver, ws in zip(sub_solvers, weights):
        for j, w in enumerate(ws):
            if w == 0.0:
                continue
            parts = [solver.make_sf_1rdm(j), solver.make_sf_2rdm(j)]
            if max_body >= 3:
                parts.append(solver.make_sf_3rdm(j))
            out = [w * x for x in parts] if out is None else [o + w * x for o, x in zip(out, parts)]
    if out is None:
        raise ValueError("make_average_rdms: no root with a non-zero weight")
    return tuple(out)
