"""TEST INFRASTRUCTURE: a host (torch CPU) stand-in for the four streaming kernels of
forte2_b200/csrc/davidson.cu, with the same call signatures as forte2_b200.davidson._Kernels.  It lets
the CPU test-suite exercise the Davidson-Liu iteration logic (and, over gloo, its sharded form) without
a GPU.  Never imported by the product."""
import numpy as np
import torch


class HostKernels:
    def __init__(self, comm=None):
        self.comm = comm
        self.launches = 0

    def _reduce(self, arr):
        t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64))
        if self.comm is not None:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.comm)
        return t.numpy().copy()

    def dots(self, A, X):
        return self._reduce((A @ X.T).numpy())

    def combine(self, A, coef, Y, beta=0.0, want_sumsq=False):
        k = Y.shape[0]
        new = beta * Y if beta != 0.0 else torch.zeros_like(Y)
        if A is not None and A.shape[0] > 0:
            c = torch.from_numpy(np.ascontiguousarray(coef, dtype=np.float64).reshape(A.shape[0], k))
            new = new + c.T @ A
        Y.copy_(new)
        return self._reduce((Y * Y).sum(dim=1).numpy()) if want_sumsq else None

    def residual(self, B, S, alpha, lam, hdiag, R, thresh):
        a = torch.from_numpy(np.ascontiguousarray(alpha, dtype=np.float64).reshape(B.shape[0], R.shape[0]))
        l = torch.from_numpy(np.ascontiguousarray(lam, dtype=np.float64))
        r = a.T @ S - l[:, None] * (a.T @ B)
        rn2 = (r * r).sum(dim=1).numpy()
        if hdiag is not None:
            d = l[:, None] - hdiag[None, :]
            r = torch.where(d.abs() > thresh, r / d, torch.zeros_like(r))
        R.copy_(r)
        return self._reduce(rn2)

    def scale(self, v, s):
        v.mul_(s)

    FUSED_MAXK, FUSED_MAXM = 4, 16
    bytes = 0

    def residual_fused(self, B, S, alpha, lam, hdiag, R, thresh):
        assert B.shape[0] <= 16 and R.shape[0] <= 4
        a = torch.from_numpy(np.ascontiguousarray(alpha, dtype=np.float64).reshape(B.shape[0], R.shape[0]))
        l = torch.from_numpy(np.ascontiguousarray(lam, dtype=np.float64))
        r = a.T @ S - l[:, None] * (a.T @ B)
        rn2 = (r * r).sum(dim=1).numpy()
        if hdiag is not None:
            d = l[:, None] - hdiag[None, :]
            r = torch.where(d.abs() > thresh, r / d, torch.zeros_like(r))
        R.copy_(r)
        return self._reduce(rn2), self._reduce((B @ R.T).numpy()), self._reduce((R @ R.T).numpy())

    def project(self, B, Y, c, T, Z):
        kin, kout = Y.shape[0], Z.shape[0]
        assert kin <= 4 and kout <= 4 and (B is None or B.shape[0] <= 16)
        y = Y.clone()
        m = 0 if B is None else B.shape[0]
        if m:
            cc = torch.from_numpy(np.ascontiguousarray(c, dtype=np.float64).reshape(m, kin))
            y = y - cc.T @ B
        TT = torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64).reshape(kin, kout))
        Z.copy_(TT.T @ y)
        d = (B @ Z.T).numpy() if m else np.zeros((0, kout))
        return self._reduce(d).reshape(m, kout), self._reduce((Z @ Z.T).numpy())



from forte2_b200.davidson import _ChunkedOps  # noqa: E402


class HostKernels16(_ChunkedOps, HostKernels):
    """The same stand-in behind the product's chunking layer (forte2_b200.davidson._ChunkedOps): the
    primitives refuse more than 16 vectors per side, exactly like f2_dl_combine / f2_dl_residual, so a solve with
    a larger subspace exercises the 16 x 16 decomposition the GPU path uses."""

    def _combine16(self, A, coef, Y, beta=0.0, want_sumsq=False):
        assert (A is None or A.shape[0] <= 16) and Y.shape[0] <= 16
        return HostKernels.combine(self, A, coef, Y, beta, want_sumsq)

    def _residual16(self, B, S, alpha, lam, hdiag, R, thresh):
        assert B.shape[0] <= 16 and R.shape[0] <= 16 and (B.shape[0] == 1 or B.stride(0) == S.stride(0))
        return HostKernels.residual(self, B, S, alpha, lam, hdiag, R, thresh)

    def dots(self, A, X):
        out = np.zeros((A.shape[0], X.shape[0]))
        for i0 in range(0, A.shape[0], 16):
            for j0 in range(0, X.shape[0], 16):
                out[i0:i0 + 16, j0:j0 + 16] = HostKernels.dots(self, A[i0:i0 + 16], X[j0:j0 + 16])
        return out
