"""Device-resident Davidson-Liu eigensolver: the GPU counterpart of forte2's
`forte2.helpers.davidsonliu.DavidsonLiuSolver` (forte2/helpers/davidsonliu.py:12-484).

Same constructor arguments, same `add_sigma_builder / add_h_diag / add_guesses / add_project_out /
solve` calls, same iteration (subspace matrix, Ritz pairs, preconditioned residuals, collapse,
iterated Gram-Schmidt with the reference's thresholds), but

* every vector lives in HBM as a contiguous row of a `(max_subspace, size)` torch CUDA tensor (the
  reference keeps `(size, max_subspace)` numpy arrays whose columns are strided,
  davidsonliu.py:75-83), so the sigma builder receives contiguous device vectors;
* every O(size) step is one streaming CUDA kernel of the C ABI (`f2_dl_dots`, `f2_dl_combine`,
  `f2_dl_residual`, `f2_dl_scale`; forte2_b200/csrc/davidson.cu); only the small subspace problem is
  solved on the host.  A kernel call takes at most 16 vectors per side; larger subspaces (many roots, e.g.
  the reference's own solver tests with 10 roots x 5 vectors) are cut into 16 x 16 pieces by `_ChunkedOps`;
* with `comm` (a torch.distributed process group) the vectors are the rank's row slab of the CI
  vector and the small reductions are all-reduced.

`sigma_builder(basis_block, sigma_block)` receives two `(m, size)` CUDA tensors (row i = vector i).
"""
from __future__ import annotations

import time
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass
class DavidsonLiuParams:
    """forte2/base_classes/params.py:107-146 (same names, defaults and validation)."""
    guess_per_root: int = 2
    ndets_per_guess: int = 10
    collapse_per_root: int = 2
    basis_per_root: int = 4
    maxiter: int = 100
    e_tol: float = 1e-12
    r_tol: float = 1e-6

    def __post_init__(self):
        if self.collapse_per_root < 1:
            raise ValueError(f"Davidson-Liu solver: collapse_per_root ({self.collapse_per_root}) must be greater "
                             "than or equal to 1.")
        if self.basis_per_root < self.collapse_per_root + 1:
            raise ValueError(f"Davidson-Liu solver: basis_per_root ({self.basis_per_root}) must be greater than or "
                             f"equal to collapse_per_root + 1 ({self.collapse_per_root + 1}).")


class _ChunkedOps:
    """combine / residual for any number of vectors on top of primitives that take at most 16 per side
    (`_combine16`, `_residual16`: one launch of f2_dl_combine / f2_dl_residual).  With at most 16 subspace
    vectors -- every large CI space: the subspace is what fills the GPU -- a call is one primitive call; beyond
    that (many roots on a small space, e.g. the reference's own solver tests with 10 roots x 5 vectors) the
    work is cut into 16 x 16 pieces, through temporaries where an output row aliases an input row."""

    MAXV = 16

    @staticmethod
    def _overlap(A, Y):
        a0, y0 = A.data_ptr(), Y.data_ptr()
        a1 = a0 + 8 * ((A.shape[0] - 1) * A.stride(0) + A.shape[1])
        y1 = y0 + 8 * ((Y.shape[0] - 1) * Y.stride(0) + Y.shape[1])
        return a0 < y1 and y0 < a1

    def combine(self, A, coef, Y, beta=0.0, want_sumsq=False):
        """Y_j = beta Y_j + sum_i coef[i, j] A_i; rows of Y may alias rows of A. Returns |Y_j|^2 or None."""
        M = self.MAXV
        k = Y.shape[0]
        p = 0 if A is None else A.shape[0]
        if p <= M and k <= M:
            return self._combine16(A, coef, Y, beta, want_sumsq)
        if p == 0:  # Y_j = beta Y_j for more than 16 rows
            sums = [self._combine16(None, np.zeros((1, min(k, j0 + M) - j0)), Y[j0:j0 + M], beta, want_sumsq)
                    for j0 in range(0, k, M)]
            return np.concatenate(sums) if want_sumsq else None
        coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(p, k)
        if k <= M and not self._overlap(A, Y):
            out = None
            for i0 in range(0, p, M):
                i1 = min(p, i0 + M)
                out = self._combine16(A[i0:i1], coef[i0:i1], Y, beta if i0 == 0 else 1.0,
                                      want_sumsq and i1 == p)
            return out
        T = Y.new_empty((k, Y.shape[1]))
        sums = []
        for j0 in range(0, k, M):
            j1 = min(k, j0 + M)
            for i0 in range(0, p, M):
                i1 = min(p, i0 + M)
                self._combine16(A[i0:i1], coef[i0:i1, j0:j1], T[j0:j1], 0.0 if i0 == 0 else 1.0, False)
        for j0 in range(0, k, M):  # every input row has been read: the outputs may now be written
            j1 = min(k, j0 + M)
            ss = self._combine16(T[j0:j1], np.eye(j1 - j0), Y[j0:j1], beta, want_sumsq)
            if want_sumsq:
                sums.append(ss)
        return np.concatenate(sums) if want_sumsq else None

    def residual(self, B, S, alpha, lam, hdiag, R, thresh):
        """R_j = preconditioned residual of Ritz pair j = sum_i alpha[i, j] (S_i - lam_j B_i); returns the squared
        norms of the residuals before preconditioning."""
        M = self.MAXV
        m, k = B.shape[0], R.shape[0]
        if m <= M and k <= M:
            return self._residual16(B, S, alpha, lam, hdiag, R, thresh)
        alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(m, k)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(k)
        out = np.zeros(k)
        if m <= M:
            for j0 in range(0, k, M):
                j1 = min(k, j0 + M)
                out[j0:j1] = self._residual16(B, S, alpha[:, j0:j1], lam[j0:j1], hdiag, R[j0:j1], thresh)
            return out
        # more than 16 basis vectors: Ritz vector and its sigma first, then the one-vector form of the kernel
        for j0 in range(0, k, M):
            j1 = min(k, j0 + M)
            X = R.new_empty((j1 - j0, R.shape[1]))
            SX = R.new_empty((j1 - j0, R.shape[1]))
            self.combine(B, alpha[:, j0:j1], X)
            self.combine(S, alpha[:, j0:j1], SX)
            for j in range(j0, j1):
                out[j] = self._residual16(X[j - j0:j - j0 + 1], SX[j - j0:j - j0 + 1], np.ones((1, 1)),
                                          lam[j:j + 1], hdiag, R[j:j + 1], thresh)[0]
        return out


class _Kernels(_ChunkedOps):
    """Thin typed wrappers over the f2_dl_* entry points for row-major (nvec, size) CUDA tensors."""

    def __init__(self, torch, device, comm=None):
        self.torch, self.device, self.comm = torch, device, comm
        self.L = _lib.lib()
        self.work = torch.empty(int(self.L.f2_dl_workspace_doubles()), dtype=torch.float64, device=device)
        self.small = torch.empty(256, dtype=torch.float64, device=device)
        self.launches = 0
        self.bytes = 0  # 8 x (vector rows read + written) x length, call by call: the HBM traffic of the passes

    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _reduce(self, t):
        if self.comm is not None:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.comm)
        return t.cpu().numpy().copy()

    @staticmethod
    def _rows(t):
        assert t.dim() == 2 and t.stride(1) == 1 and t.dtype.is_floating_point and t.element_size() == 8
        return t.data_ptr(), (t.stride(0) if t.shape[0] > 1 else t.shape[1]), t.shape[0], t.shape[1]

    def dots(self, A, X):
        """(p, q) numpy matrix of <A_i | X_j>, summed over ranks."""
        ap, lda, p, n = self._rows(A)
        xp, ldx, q, n2 = self._rows(X)
        assert n == n2
        out = np.zeros((p, q))
        for i0 in range(0, p, 16):
            pc = min(16, p - i0)
            for j0 in range(0, q, 16):
                qc = min(16, q - j0)
                res = self.small[: pc * qc]
                _lib.check(self.L.f2_dl_dots(ap + 8 * lda * i0, lda, pc, xp + 8 * ldx * j0, ldx, qc, n,
                                             res.data_ptr(), self.work.data_ptr(), self._stream()))
                self.launches += 2 * ((qc + 3) // 4)
                self.bytes += 8 * n * (pc * ((qc + 3) // 4) + qc)
                out[i0:i0 + pc, j0:j0 + qc] = self._reduce(res).reshape(pc, qc)
        return out

    def _combine16(self, A, coef, Y, beta=0.0, want_sumsq=False):
        """one f2_dl_combine launch (at most 16 input and 16 output rows; outputs may alias inputs)"""
        yp, ldy, k, n = self._rows(Y)
        if A is None or A.shape[0] == 0:
            ap, lda, p = yp, ldy, 0
            coef = np.zeros((1, k))
        else:
            ap, lda, p, n2 = self._rows(A)
            assert n == n2
        assert p <= 16 and k <= 16
        coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(max(p, 1), k)
        ss = self.small[:k] if want_sumsq else None
        _lib.check(self.L.f2_dl_combine(ap, lda, p, coef.ctypes.data_as(_lib.f64p), yp, ldy, k, float(beta), n,
                                        ss.data_ptr() if want_sumsq else None, self.work.data_ptr(),
                                        self._stream()))
        self.launches += 2 if want_sumsq else 1
        self.bytes += 8 * n * (p + k + (k if beta != 0.0 else 0))
        return self._reduce(ss) if want_sumsq else None

    def _residual16(self, B, S, alpha, lam, hdiag, R, thresh):
        """one f2_dl_residual launch (at most 16 basis vectors and 16 roots)"""
        bp, ld, m, n = self._rows(B)
        sp, lds, m2, _ = self._rows(S)
        rp, ldr, k, _ = self._rows(R)
        assert m == m2 and (ld == lds or m == 1) and m <= 16 and k <= 16
        alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(m, k)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(k)
        out = self.small[:k]
        _lib.check(self.L.f2_dl_residual(bp, sp, ld, m, alpha.ctypes.data_as(_lib.f64p),
                                         lam.ctypes.data_as(_lib.f64p),
                                         hdiag.data_ptr() if hdiag is not None else None, rp, ldr, k,
                                         float(thresh), n, out.data_ptr(), self.work.data_ptr(), self._stream()))
        self.launches += 2
        self.bytes += 8 * n * (2 * m + k + 1)
        return self._reduce(out)

    def scale(self, v, s):
        assert v.dim() == 1 and v.stride(0) == 1
        _lib.check(self.L.f2_dl_scale(v.data_ptr(), float(s), v.shape[0], self._stream()))
        self.launches += 1
        self.bytes += 16 * v.shape[0]

    # ---- fused passes (at most 4 roots, 16 basis vectors): f2_dl_residual_fused / f2_dl_project
    FUSED_MAXK, FUSED_MAXM = 4, 16

    def residual_fused(self, B, S, alpha, lam, hdiag, R, thresh):
        """R_j = preconditioned residual of Ritz pair j; returns (|r_j|^2 before preconditioning (k),
        <B_i|R_j> (m, k), <R_j|R_j'> (k, k)), summed over ranks"""
        bp, ld, m, n = self._rows(B)
        sp, lds, m2, _ = self._rows(S)
        rp, ldr, k, _ = self._rows(R)
        assert m == m2 and (ld == lds or m == 1) and m <= self.FUSED_MAXM and k <= self.FUSED_MAXK
        alpha = np.ascontiguousarray(alpha, dtype=np.float64).reshape(m, k)
        lam = np.ascontiguousarray(lam, dtype=np.float64).reshape(k)
        nout = k + m * k + k * k
        out = self.small[:nout]
        _lib.check(self.L.f2_dl_residual_fused(bp, sp, ld, m, alpha.ctypes.data_as(_lib.f64p),
                                               lam.ctypes.data_as(_lib.f64p),
                                               hdiag.data_ptr() if hdiag is not None else None, rp, ldr, k,
                                               float(thresh), n, out.data_ptr(), self.work.data_ptr(), self._stream()))
        self.launches += 2
        self.bytes += 8 * n * (2 * m + k + 1)
        o = self._reduce(out)
        return o[:k], o[k:k + m * k].reshape(m, k), o[k + m * k:].reshape(k, k)

    def project(self, B, Y, c, T, Z):
        """Z_j = sum_l T[l, j] (Y_l - sum_i c[i, l] B_i); Z may alias Y.  Returns (<B_i|Z_j> (m, kout),
        <Z_j|Z_j'> (kout, kout)), summed over ranks"""
        yp, ldy, kin, n = self._rows(Y)
        zp, ldz, kout, _ = self._rows(Z)
        if B is None or B.shape[0] == 0:
            bp, ld, m = yp, ldy, 0
            c = np.zeros((1, kin))
        else:
            bp, ld, m, _ = self._rows(B)
        assert m <= self.FUSED_MAXM and kin <= self.FUSED_MAXK and kout <= self.FUSED_MAXK
        c = np.ascontiguousarray(c, dtype=np.float64).reshape(max(m, 1), kin)
        T = np.ascontiguousarray(T, dtype=np.float64).reshape(kin, kout)
        nout = m * kout + kout * kout
        out = self.small[:nout]
        _lib.check(self.L.f2_dl_project(bp, ld, m, yp, ldy, kin, c.ctypes.data_as(_lib.f64p),
                                        T.ctypes.data_as(_lib.f64p), zp, ldz, kout, n, out.data_ptr(),
                                        self.work.data_ptr(), self._stream()))
        self.launches += 2
        self.bytes += 8 * n * (m + kin + kout)
        o = self._reduce(out)
        return o[:m * kout].reshape(m, kout), o[m * kout:].reshape(kout, kout)


class DavidsonLiuSolver:
    """Davidson-Liu on device vectors. See the module docstring; reference: davidsonliu.py:12-484."""

    def __init__(self, size, nroot, davidson_liu_params=None, energy_shift=None, log_level=3, dtype=np.float64,
                 device=None, comm=None, global_size=None, log=None, _kernels=None):
        import torch
        self.torch = torch
        self.size, self.nroot = int(size), int(nroot)
        self.davidson_liu_params = davidson_liu_params or DavidsonLiuParams()
        self.energy_shift, self.log_level, self.comm = energy_shift, log_level, comm
        self.global_size = int(global_size) if global_size is not None else self.size
        if np.dtype(dtype) != np.float64:
            raise TypeError("the device Davidson-Liu solver is float64 only")
        p = self.davidson_liu_params
        self.collapse_per_root, self.basis_per_root = p.collapse_per_root, p.basis_per_root
        self.maxiter, self.e_tol, self.r_tol = p.maxiter, p.e_tol, p.r_tol
        if self.global_size <= 0:
            raise ValueError("Davidson-Liu solver called with space of dimension smaller than 1.")
        if self.nroot <= 0:
            raise ValueError("Davidson-Liu solver called with zero roots.")
        self.collapse_size = min(self.collapse_per_root * self.nroot, self.global_size)
        self.max_subspace_size = min(self.basis_per_root * self.nroot, self.global_size)
        # _kernels: test hook (tests/ inject a host stand-in for the four streaming kernels to exercise the
        # iteration logic without a GPU); the product path below has no CPU fallback
        if _kernels is None:
            if not torch.cuda.is_available():
                raise RuntimeError("forte2_b200: no CUDA device available (the Davidson-Liu solver has no CPU "
                                   "fallback)")
            if device is None:
                device = torch.cuda.current_device()
            self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        else:
            self.device = torch.device(device if device is not None else "cpu")
        kw = dict(dtype=torch.float64, device=self.device)
        self.b = torch.zeros((self.max_subspace_size, self.size), **kw)
        self.sigma = torch.zeros((self.max_subspace_size, self.size), **kw)
        self.r = torch.zeros((self.nroot, self.size), **kw)
        self.h_diag = None
        self.G = np.zeros((self.max_subspace_size, self.max_subspace_size))
        self.alpha = np.zeros_like(self.G)
        self.lam = np.zeros(self.max_subspace_size)
        self.lam_old = np.zeros_like(self.lam)
        self.schmidt_discard_threshold = 1e-8
        self.schmidt_orthogonality_threshold = 1e-12
        self.preconditioner_denom_threshold = 1e-8
        self.basis_size = 0
        self.sigma_size = 0
        self._build_sigma = None
        self._executed = False
        self._guesses = None
        self._proj_out = None
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(20261017)
        self.k = _kernels if _kernels is not None else _Kernels(torch, self.device, comm)
        self._log = log
        self.iter_times = []
        self.history = []

    # ---- reference interface ------------------------------------------------------------------
    def add_sigma_builder(self, sigma_builder):
        self._build_sigma = sigma_builder

    def add_h_diag(self, h_diag):
        t = self._as_device(h_diag)
        if t.shape != (self.size,):
            raise ValueError("h_diag must be shape (size,)")
        self.h_diag = t.contiguous()

    def add_guesses(self, guesses):
        """numpy (size, n_guess) as the reference takes them, or a CUDA tensor (n_guess, size)."""
        torch = self.torch
        if isinstance(guesses, np.ndarray):
            if guesses.ndim != 2 or guesses.shape[0] != self.size:
                raise ValueError("guesses must be shape (size, n_guess)")
            g = torch.from_numpy(np.ascontiguousarray(guesses.T, dtype=np.float64)).to(self.device)
        else:
            g = guesses
            if g.dim() != 2 or g.shape[1] != self.size:
                raise ValueError("device guesses must be shape (n_guess, size)")
        if not 1 <= g.shape[0] <= self.max_subspace_size:
            raise ValueError("between 1 and max_subspace_size guesses are needed")
        self._guesses = g

    def add_project_out(self, project_out):
        """Vectors to project out of sigma and of the corrections (davidsonliu.py:147-182); each may be
        shorter than `size` (zero padded).  Orthonormalised on the host like the reference (QR)."""
        if len(project_out) == 0:
            return
        if self.comm is not None:
            raise NotImplementedError("project_out is not supported with sharded vectors")
        arrs = [np.asarray(v.cpu() if hasattr(v, "cpu") else v, dtype=np.float64) for v in project_out]
        for v in arrs:
            if len(v) > self.size:
                raise ValueError(f"Each project_out vector must have length at most {self.size}, but got {len(v)}")
        Q, _ = np.linalg.qr(np.stack(arrs, axis=1), mode="reduced")
        self._proj_out = self.torch.from_numpy(np.ascontiguousarray(Q.T)).to(self.device)

    def add_project_out_sparse(self, indices, vectors):
        """Vectors to project out that live on a few components only: vectors[j] holds the components of vector j on
        the (local) vector indices `indices`, zero elsewhere (the wrong-spin combinations of the guess determinants,
        forte2/sci/sci.py:437-441).  On one GPU they are orthonormalised here (QR, like add_project_out); with
        sharded vectors the caller passes every rank its components of globally orthonormal vectors."""
        torch = self.torch
        V = np.asarray(vectors, dtype=np.float64)
        if V.ndim != 2 or V.shape[1] != len(indices):
            raise ValueError("vectors must be shape (n_vectors, len(indices))")
        if V.shape[0] == 0:
            return
        if self.comm is None:
            Q, _ = np.linalg.qr(V.T, mode="reduced")
            V = np.ascontiguousarray(Q.T)
        self._proj_sparse = (torch.from_numpy(np.ascontiguousarray(indices, dtype=np.int64)).to(self.device),
                             torch.from_numpy(np.ascontiguousarray(V)).to(self.device))

    def _do_project_out_sparse(self, rows):
        idx, Q = self._proj_sparse
        sub = rows[:, idx]                    # (k, nidx)
        v = sub @ Q.T                         # overlaps (k, nproj)
        if self.comm is not None:
            import torch.distributed as dist
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=self.comm)
        rows[:, idx] = sub - v @ Q

    def do_project_out(self, rows):
        if rows.shape[0] > 0 and getattr(self, "_proj_sparse", None) is not None:
            self._do_project_out_sparse(rows)
        if self._proj_out is None or rows.shape[0] == 0:
            return
        psize = self._proj_out.shape[1]
        sub = rows[:, :psize]
        for u in self._proj_out:
            u2 = u.unsqueeze(0)
            for j in range(rows.shape[0]):
                tgt = sub[j:j + 1]
                c = self.k.dots(u2, tgt)
                self.k.combine(u2, -c, tgt, beta=1.0)

    # ---- helpers ------------------------------------------------------------------------------
    def _as_device(self, x):
        torch = self.torch
        if isinstance(x, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(self.device)
        return x.to(self.device, dtype=torch.float64)

    def _say(self, msg):
        if self._log is not None:
            self._log(msg)

    def _add_and_orthonormalize(self, cand_rows):
        """davidsonliu.py:415-484: candidates (rows) are orthonormalised one by one against
        b[:basis_size] and the rows accepted so far, and appended to b.  Returns the number added.
        Classical Gram-Schmidt with re-orthogonalisation: each cycle is one f2_dl_dots and one
        f2_dl_combine pass; the normalisation is carried as a pending factor and folded into the
        next pass."""
        added = 0
        self._accepted = [False] * cand_rows.shape[0]
        for j in range(cand_rows.shape[0]):
            cur = self.basis_size + added
            if cur >= self.max_subspace_size:
                break
            slot = self.b[cur:cur + 1]
            slot.copy_(cand_rows[j:j + 1])
            basis = self.b[:cur]
            s = 1.0        # the candidate is s * slot
            ok = False
            normalised = False
            for cycle in range(10):
                if cur > 0:
                    c = self.k.dots(basis, slot)[:, 0] * s
                else:
                    c = np.zeros(0)
                if normalised and (cur == 0 or np.abs(c).max() < self.schmidt_orthogonality_threshold):
                    ok = True
                    break
                ss = self.k.combine(basis if cur > 0 else None, -c.reshape(-1, 1), slot, beta=s, want_sumsq=True)
                normv = float(np.sqrt(ss[0]))
                if not np.isfinite(normv) or normv < self.schmidt_discard_threshold:
                    break
                s = 1.0 / normv
                normalised = True
            if ok:
                if s != 1.0:
                    self.k.scale(slot[0], s)
                added += 1
                self._accepted[j] = True
        return added

    def _fused_ok(self, m, k):
        return hasattr(self.k, "project") and 0 < k <= self.k.FUSED_MAXK and m <= self.k.FUSED_MAXM

    def _projecting(self):
        return self._proj_out is not None or getattr(self, "_proj_sparse", None) is not None

    @staticmethod
    def _gram_schmidt_coefficients(g, alive, room, discard):
        """Sequential Gram-Schmidt of the columns y_j (Gram matrix g) in coefficient form: returns T (k, nacc) with
        z = y T orthonormal, skipping columns whose remainder is shorter than `discard` (davidsonliu.py:460-463)
        and stopping at `room` accepted columns, and the list of accepted columns."""
        k = g.shape[0]
        cols, acc = [], []
        for j in range(k):
            if not alive[j] or len(cols) >= room:
                continue
            t = np.zeros(k)
            t[j] = 1.0
            for _ in range(2):
                for ta in cols:
                    t = t - (ta @ g @ t) * ta
            nrm2 = float(t @ g @ t)
            if not np.isfinite(nrm2) or nrm2 < discard * discard:
                continue
            cols.append(t / np.sqrt(nrm2))
            acc.append(j)
        T = np.stack(cols, axis=1) if cols else np.zeros((k, 0))
        return T, acc

    def _add_block_orthonormalized(self, cand, c):
        """Block form of _add_and_orthonormalize for at most 4 candidates and 16 basis vectors (the usual CI
        solve): cand (k, size) is projected against b[:basis_size] with the overlaps c = <b_i|cand_j> already
        known from the pass that produced it (None: computed here), re-projected until the overlaps of the
        normalised remainders are below the reference's threshold, then orthonormalised among themselves and
        written to the next basis slots.  Every pass reads the basis once for all candidates and returns the
        overlaps of its own result.  Returns the number of vectors added, or -1 if the block did not reach the
        thresholds (the caller then takes the vector-by-vector path)."""
        k, cur = cand.shape[0], self.basis_size
        room = self.max_subspace_size - cur
        self._accepted = [False] * k
        if room <= 0:
            return 0
        B = self.b[:cur] if cur > 0 else None
        eye = np.eye(k)
        if c is None:
            c = self.k.dots(B, cand) if cur > 0 else np.zeros((0, k))
        tiny = self.schmidt_orthogonality_threshold
        d = g = None
        for cycle in range(10):
            d, g = self.k.project(B, cand, c, eye, cand)
            norms = np.sqrt(np.maximum(np.diag(g), 0.0))
            alive = np.isfinite(norms) & (norms >= self.schmidt_discard_threshold)
            if not alive.any():
                return 0
            rel = max((np.abs(d[:, j]).max() / norms[j] for j in range(k) if alive[j] and cur > 0), default=0.0)
            if rel < tiny:
                break
            c = d
        else:
            return -1
        # remainder after one more (tiny) projection has the Gram matrix g - d^T d to first order
        g2 = g - d.T @ d if cur > 0 else g
        T, acc = self._gram_schmidt_coefficients(g2, alive, room, self.schmidt_discard_threshold)
        if not acc:
            return 0
        slots = self.b[cur:cur + len(acc)]
        dz, gz = self.k.project(B, cand, d if cur > 0 else np.zeros((0, k)), T, slots)
        for refine in range(3):
            off = max(np.abs(dz).max() if cur > 0 else 0.0, np.abs(gz - np.eye(len(acc))).max())
            if off < tiny:
                break
            # nearly dependent candidates: the Gram form loses digits; one more pass on the result restores them
            T2, acc2 = self._gram_schmidt_coefficients(gz - (dz.T @ dz if cur > 0 else 0.0), np.ones(len(acc), bool),
                                                       len(acc), self.schmidt_discard_threshold)
            if len(acc2) != len(acc):
                return -1
            dz, gz = self.k.project(B, slots, dz if cur > 0 else np.zeros((0, len(acc))), T2, slots)
        else:
            return -1
        for j in acc:
            self._accepted[j] = True
        return len(acc)

    def _events(self):
        if self.device.type == "cuda":
            return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

        class _Wall:  # host stand-in used with the test hook
            def record(self):
                self.t = time.perf_counter()

            def synchronize(self):
                pass

            def elapsed_time(self, other):
                return (other.t - self.t) * 1e3
        return _Wall(), _Wall()

    def _random_rows(self, n):
        t = self.torch.rand((n, self.size), dtype=self.torch.float64, device=self.device, generator=self._gen)
        return t * 2.0 - 1.0

    # ---- the solver ---------------------------------------------------------------------------
    def solve(self):
        torch = self.torch
        if self._build_sigma is None:
            raise ValueError("Sigma builder function is not set.")
        if self.h_diag is None:
            raise ValueError("Hamiltonian diagonal is not set.")
        nroot = self.nroot

        if not self._executed:
            G = self._guesses if self._guesses is not None else self._random_rows(nroot)
            G = G.clone()
            self.do_project_out(G)
            self.basis_size = 0
            self.sigma_size = 0
            self.basis_size = self._add_and_orthonormalize(G)
            if self.basis_size == 0:
                raise ValueError("no linearly independent guess vector")
            self.lam_old[:] = 0.0
            self.G[:] = 0.0
        self.iter = 0
        self.converged = False
        self.iter_times = []
        self.iter_bytes = []  # bytes moved by the solver's own passes in each iteration (sigma builds excluded)
        self.history = []
        g_valid = self.sigma_size  # rows/columns of self.G that are up to date

        for self.iter in range(self.maxiter):
            t0 = time.perf_counter()
            bytes0 = getattr(self.k, "bytes", 0)
            ev0, ev1 = self._events()
            ev0.record()
            # 2. sigma for the new basis vectors
            m_new = self.basis_size - self.sigma_size
            if m_new > 0:
                Bblock = self.b[self.sigma_size:self.basis_size]
                Sblock = self.sigma[self.sigma_size:self.basis_size]
                self._build_sigma(Bblock, Sblock)
                self.do_project_out(Sblock)
                self.sigma_size = self.basis_size
            m = self.basis_size
            Bblk, Sblk = self.b[:m], self.sigma[:m]
            # 3. subspace matrix: only the new columns are computed (the reference recomputes B^T S)
            if g_valid < m:
                cols = self.k.dots(Bblk, self.sigma[g_valid:m])  # (m, m - g_valid)
                self.G[:m, g_valid:m] = cols
                self.G[g_valid:m, :g_valid] = cols[:g_valid].T
                g_valid = m
            Gm = 0.5 * (self.G[:m, :m] + self.G[:m, :m].T)
            lam, alpha = np.linalg.eigh(Gm)
            if self.energy_shift is not None:
                idx = np.argsort(np.abs(lam - self.energy_shift))
                lam, alpha = lam[idx], alpha[:, idx]
            self.lam[:m] = lam
            self.alpha[:m, :m] = alpha
            kr = min(nroot, m)
            ar, lamr = alpha[:, :kr], lam[:kr]

            # 4. residuals, their norms and the preconditioned corrections in one pass
            fused = self._fused_ok(m, kr)
            c_res = None
            if fused:
                rn2, c_res, _ = self.k.residual_fused(Bblk, Sblk, ar, lamr, self.h_diag, self.r[:kr],
                                                      self.preconditioner_denom_threshold)
            else:
                rn2 = self.k.residual(Bblk, Sblk, ar, lamr, self.h_diag, self.r[:kr],
                                      self.preconditioner_denom_threshold)
            rnorms = np.sqrt(rn2)
            avg_e = float(lamr.mean())
            max_de = float(np.max(np.abs(lamr - self.lam_old[:kr])))
            max_r = float(rnorms.max())
            conv = (max_de < self.e_tol and max_r < self.r_tol) or (m == self.global_size)
            if conv or self.iter == self.maxiter - 1:
                self.converged = bool(conv)
                ev1.record()
                ev1.synchronize()
                self.iter_times.append(ev0.elapsed_time(ev1) * 1e-3)
                self.iter_bytes.append(getattr(self.k, "bytes", 0) - bytes0)
                self.history.append((self.iter, avg_e, max_de, max_r, m, m_new))
                break
            # Root locking (a deliberate difference from davidsonliu.py:306-366): the reference adds a
            # correction for every root and replaces a discarded one (norm < 1e-8 after projection, i.e. a
            # root converged far below r_tol) by a random vector, which costs a sigma build and perturbs the
            # other roots.  Here only roots that are not converged yet get a correction, the subspace is
            # collapsed only when those do not fit, and a random vector replaces a discarded correction of
            # an unconverged root only.  With every root active this is the reference's iteration.
            de = np.abs(lamr - self.lam_old[:kr])
            active = [j for j in range(kr) if rnorms[j] >= self.r_tol or de[j] >= self.e_tol]
            self.lam_old[:kr] = lamr

            # 5. collapse
            if self.basis_size + len(active) > self.max_subspace_size:
                k = min(self.collapse_size, self.basis_size)
                ak = alpha[:, :k]
                self.k.combine(Bblk, ak, self.b[:k])
                self.k.combine(Sblk, ak, self.sigma[:k])
                Gk = ak.T @ Gm @ ak
                self.G[:] = 0.0
                self.G[:k, :k] = Gk
                self.basis_size = self.sigma_size = g_valid = k
                m = k
                if c_res is not None:  # overlaps with the collapsed basis b_new = B alpha_k
                    c_res = ak.T @ c_res

            # 6. corrections: project out, orthonormalise against the basis, append
            to_add = min(len(active), self.max_subspace_size - self.basis_size)
            if to_add == 0:
                self.converged = True
                break
            if active[:to_add] != list(range(to_add)):
                for dst, src in enumerate(active[:to_add]):  # dst <= src: compact the active rows
                    if dst != src:
                        self.r[dst].copy_(self.r[src])
            R0 = self.r[:to_add]
            self.do_project_out(R0)
            added = -1
            if fused and c_res is not None:
                # after a project-out the overlaps returned by the residual pass are stale: recomputed inside
                added = self._add_block_orthonormalized(R0, None if self._projecting() else c_res[:, active[:to_add]])
            if added < 0:
                added = self._add_and_orthonormalize(R0)
            self.basis_size += added
            msg = ""
            missing = to_add - added
            if missing > 0:
                temp = self._random_rows(missing)
                self.do_project_out(temp)
                added2 = self._add_and_orthonormalize(temp)
                self.basis_size += added2
                msg = f"+{added2} rand"
            ev1.record()
            ev1.synchronize()
            self.iter_times.append(ev0.elapsed_time(ev1) * 1e-3)
            self.iter_bytes.append(getattr(self.k, "bytes", 0) - bytes0)
            self.history.append((self.iter, avg_e, max_de, max_r, self.basis_size, m_new))
            self._say(f"{self.iter:4d} {avg_e:20.12f} {max_de:+12.4e} {max_r:+12.4e} {self.basis_size:4d} "
                      f"{time.perf_counter() - t0:10.4f} {msg}")

        # final Ritz vectors, in place in the first nroot rows of b
        m = self.basis_size
        kr = min(nroot, m)
        lamr = self.lam[:kr].copy()
        self.k.combine(self.b[:m], self.alpha[:m, :kr], self.b[:kr])
        evecs = self.b[:kr]
        self.basis_size = kr
        self.sigma_size = 0
        self._executed = True
        return lamr, evecs
