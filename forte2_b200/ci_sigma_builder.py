"""CISigmaBuilder: the Python face of forte2.lib.ci_helpers.CISigmaBuilder
(forte2/api/ci_helpers_api.cc:88-109,146-148; hot-path subset) running on one B200.

`Hamiltonian(basis, sigma)` accepts
  * torch CUDA float64 tensors (or any DLPack-exporting device array): zero-copy, asynchronous on the
    current torch stream;
  * numpy float64 vectors, as the reference does: staged through pinned memory inside the C ABI
    (f2_sigma_apply_host), synchronous.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .ci_strings import CIStrings


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch")


class CISigmaBuilder:
    def __init__(self, lists: CIStrings, E: float, H, V, log_level: int = 3, device: int | None = None):
        if not isinstance(lists, CIStrings):
            raise TypeError("lists must be a forte2_b200.CIStrings")
        self.ci_strings = lists  # keep-alive (the reference leaves this to the caller, ci.py:172)
        n = lists.norb
        H = np.asarray(H, dtype=np.float64)
        V = np.asarray(V, dtype=np.float64)
        if H.ndim != 2:
            raise RuntimeError("H must be a 2D matrix.")
        if H.shape != (n, n):
            raise RuntimeError("H shape does not match the number of orbitals.")
        if V.ndim != 4:
            raise RuntimeError("V must be a 4D tensor.")
        if V.shape != (n, n, n, n):
            raise RuntimeError("V shape does not match the number of orbitals.")
        H = np.ascontiguousarray(H)
        V = np.ascontiguousarray(V)
        if device is None:
            device = self._current_device()
        self.device = int(device)
        self._ints = (float(E), H, V)  # the reference keeps handles of H and V too (ci_sigma_builder.h:294-296)
        self._memory_bytes = 1073741824  # ci_sigma_builder.h:300
        L = _lib.lib()
        self._h = L.f2_sigma_create(lists._h, float(E), H.ctypes.data_as(_lib.f64p),
                                    V.ctypes.data_as(_lib.f64p), int(log_level), self.device)
        if not self._h:
            raise RuntimeError(_lib.last_error())

    @staticmethod
    def _current_device():
        try:
            import torch
            if torch.cuda.is_available():
                return torch.cuda.current_device()
        except ImportError:
            pass
        return 0

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.f2_sigma_destroy(h)

    def set_Hamiltonian(self, E, H, V):
        n = self.ci_strings.norb
        H = np.ascontiguousarray(H, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        if H.shape != (n, n):
            raise RuntimeError("H shape does not match the number of orbitals.")
        if V.shape != (n, n, n, n):
            raise RuntimeError("V shape does not match the number of orbitals.")
        _lib.check(_lib.lib().f2_sigma_set_hamiltonian(self._h, float(E), H.ctypes.data_as(_lib.f64p),
                                                       V.ctypes.data_as(_lib.f64p)))
        self._ints = (float(E), H, V)

    def set_algorithm(self, algorithm: str):
        _lib.check(_lib.lib().f2_sigma_set_algorithm(self._h, str(algorithm).encode()))

    def get_algorithm(self) -> str:
        return _lib.lib().f2_sigma_get_algorithm(self._h).decode()

    def set_memory(self, memory: int):
        _lib.check(_lib.lib().f2_sigma_set_memory(self._h, int(memory)))
        self._memory_bytes = int(memory) * 1024 * 1024

    def set_log_level(self, level: int):
        _lib.check(_lib.lib().f2_sigma_set_log_level(self._h, int(level)))

    def Hamiltonian(self, basis, sigma) -> None:
        L = _lib.lib()
        if isinstance(basis, np.ndarray) and isinstance(sigma, np.ndarray):
            if basis.dtype != np.float64 or sigma.dtype != np.float64:
                raise TypeError("basis and sigma must be float64")
            if basis.ndim != 1 or sigma.ndim != 1:
                raise TypeError("basis and sigma must be one-dimensional")
            if (basis.size > 1 and basis.strides[0] != 8) or (sigma.size > 1 and sigma.strides[0] != 8):
                raise RuntimeError("np_vector must have stride 1 to convert to std::span.")
            if basis.shape[0] != sigma.shape[0]:
                raise ValueError("basis and sigma must have the same length")
            _lib.check(L.f2_sigma_apply_host(self._h, basis.ctypes.data_as(_lib.f64p),
                                             sigma.ctypes.data_as(_lib.f64p), basis.shape[0]))
            return
        import torch
        if not _is_torch_tensor(basis):
            basis = torch.from_dlpack(basis)
        if not _is_torch_tensor(sigma):
            sigma = torch.from_dlpack(sigma)
        if basis.dim() == 2 and sigma.dim() == 2:
            # a block of trial vectors, row v = vector v (the Davidson-Liu basis block): one call for all of them
            for name, t in (("basis", basis), ("sigma", sigma)):
                if not t.is_cuda or t.dtype != torch.float64:
                    raise TypeError(f"{name} must be a float64 CUDA tensor")
                if t.shape[1] > 1 and t.stride(1) != 1:
                    raise RuntimeError("np_vector must have stride 1 to convert to std::span.")
                if t.device.index != self.device:
                    raise ValueError(f"{name} lives on cuda:{t.device.index}, the builder on cuda:{self.device}")
            if basis.shape != sigma.shape:
                raise ValueError("basis and sigma must have the same shape")
            nvec, n = basis.shape
            ld = basis.stride(0) if nvec > 1 else n
            if nvec > 1 and sigma.stride(0) != ld:
                raise ValueError("basis and sigma blocks must have the same row stride")
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(L.f2_sigma_apply_batch(self._h, basis.data_ptr(), sigma.data_ptr(), n, nvec, ld, stream))
            return
        for name, t in (("basis", basis), ("sigma", sigma)):
            if not t.is_cuda or t.dtype != torch.float64 or t.dim() != 1:
                raise TypeError(f"{name} must be a one-dimensional float64 CUDA tensor")
            if t.numel() > 1 and t.stride(0) != 1:
                raise RuntimeError("np_vector must have stride 1 to convert to std::span.")
            if t.device.index != self.device:
                raise ValueError(f"{name} lives on cuda:{t.device.index}, the builder on cuda:{self.device}")
        if basis.numel() != sigma.numel():
            raise ValueError("basis and sigma must have the same length")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(L.f2_sigma_apply(self._h, basis.data_ptr(), sigma.data_ptr(), basis.numel(), stream))

    def avg_build_time(self):
        out = np.zeros(4)
        _lib.check(_lib.lib().f2_sigma_times(self._h, out.ctypes.data_as(_lib.f64p)))
        return [float(x) for x in out]

    def last_build_ms(self):
        """[H0, same-spin alpha, same-spin beta, opposite spin, total] of the last build (device ms)."""
        out = np.zeros(5)
        _lib.check(_lib.lib().f2_sigma_last_ms(self._h, out.ctypes.data_as(_lib.f64p)))
        return [float(x) for x in out]

    def launch_count(self) -> int:
        return int(_lib.lib().f2_sigma_launch_count(self._h))

    def t3_shape(self):
        """dict describing the contraction the opposite-spin kernel executes (see the header)."""
        out = np.zeros(6, dtype=np.int64)
        _lib.check(_lib.lib().f2_sigma_t3_shape(self._h, out.ctypes.data_as(_lib.i64p)))
        return {"M": int(out[0]), "M_pad": int(out[1]), "K_max": int(out[2]), "k_cols": int(out[3]),
                "pair_compressed": bool(out[4]), "form": "gather" if out[5] else "scatter"}

    def form_Hdiag_det_part(self, rank: int, world: int, local_size: int, spin_averaged: bool = False):
        """determinant energies of the alpha rows rank `rank` of `world` owns, in the local layout of
        forte2_b200/dist.py (f2_sigma_hdiag_part / f2_sigma_hdiag_spin_averaged)"""
        import torch
        out = torch.empty(local_size, dtype=torch.float64, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        fn = _lib.lib().f2_sigma_hdiag_spin_averaged if spin_averaged else _lib.lib().f2_sigma_hdiag_part
        _lib.check(fn(self._h, out.data_ptr(), rank, world, stream))
        return out

    def form_Hdiag_csf(self, dets, spin_adapter, spin_adapt_full_preconditioner: bool = False):
        """Davidson preconditioner in the CSF basis (CISigmaBuilder::form_Hdiag_csf,
        forte2/ci/ci_sigma_builder.cc:317-370) as a CUDA tensor of ncsf doubles.  `dets` is accepted for
        signature compatibility and not used (the adapter knows its determinants).  The default form
        H_II ~ sum_d c_dI^2 <d|H|d> is computed on the device; the exact form <I|H|I> (which the reference itself
        calls too slow to use, ci_sigma_builder.cc:326-332) is computed on the host with Slater rules, as the
        reference does, and copied to the device."""
        import torch
        if spin_adapter.ci_strings is not self.ci_strings and spin_adapter.ndet != self.ci_strings.ndet:
            raise ValueError("the spin adapter was prepared for a different determinant space")
        if spin_adapt_full_preconditioner:
            return torch.from_numpy(exact_hdiag_csf(self.ci_strings, *self._ints, spin_adapter)).to(
                f"cuda:{self.device}")
        e_det = self.form_Hdiag_det()
        out = torch.empty(spin_adapter.ncsf, dtype=torch.float64, device=e_det.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().f2_adapter_hdiag_csf(spin_adapter._h, e_det.data_ptr(), out.data_ptr(), self.device,
                                                   stream))
        return out

    # ---- host-side helpers of the callers around sigma (guess vectors, exact diagonalisation of small spaces):
    # Slater rules over the few determinants involved, f2_slater_matrix; no device work
    def slater_rules_csf(self, dets, spin_adapter, I: int, J: int) -> float:
        """<I|H|J> between two CSFs (CISigmaBuilder::slater_rules_csf, forte2/ci/ci_sigma_builder.cc:278-297;
        the guess Hamiltonian of forte2/ci/ci.py:405-441).  `dets` is accepted for signature compatibility;
        `spin_adapter` needs csf(i) -> [(determinant address, coefficient)], which both this package's and
        the reference's adapter offer."""
        return csf_matrix_element(self.ci_strings, *self._ints, spin_adapter, I, J)

    def energy_csf(self, dets, spin_adapter, I: int) -> float:
        """<I|H|I> of one CSF (CISigmaBuilder::energy_csf, ci_sigma_builder.cc:299-315)"""
        return self.slater_rules_csf(dets, spin_adapter, I, I)

    def form_H_csf(self, dets, spin_adapter) -> np.ndarray:
        """Dense Hamiltonian in the CSF basis, (ncsf, ncsf) numpy array (CISigmaBuilder::form_H_csf,
        ci_sigma_builder.cc:372-392; the exact-diagonalisation path of forte2/ci/ci.py:305-318), with the
        reference's memory check and message.  T^t H_det T over the determinants the CSFs use."""
        ncsf = _ncsf(spin_adapter)
        if ncsf * ncsf * 8 > self._memory_bytes:
            raise RuntimeError("Not enough memory to form the full Hamiltonian matrix in the CSF basis. "
                               "Increase the memory limit or use a different algorithm.")
        return csf_hamiltonian(self.ci_strings, *self._ints, spin_adapter)

    # ---- reduced density matrices (forte2/api/ci_helpers_api.cc:110-130): device kernels for the spin blocks
    # (f2_sigma_rdm, csrc/rdm_kernels.cuh), the spin-free combinations assembled on the host like the reference
    def _rdm(self, kind, C_left, C_right, shape, part=None, device_out=False):
        """part = (i, n): only the contributions of slice i of n of the hole strings (f2_sigma_rdm_part; the parts add up
        to the block); device_out: the result stays on the device as a torch tensor (for an all-reduce)"""
        import torch
        dev = torch.device("cuda", self.device)
        ndet = self.ci_strings.ndet

        def vec(x, name):
            t = x if _is_torch_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
            if not t.is_cuda or t.dtype != torch.float64 or t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
                raise TypeError(f"{name} must be a one-dimensional contiguous float64 vector")
            if t.numel() != ndet:
                raise RuntimeError("Size of CI vector does not match the number of determinants.")
            return t

        left, right = vec(C_left, "C_left"), vec(C_right, "C_right")
        count = int(np.prod(shape))
        if count == 0 or ndet == 0:
            return torch.zeros(shape, dtype=torch.float64, device=dev) if device_out else np.zeros(shape)
        out = torch.empty(count, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if part is None:
            _lib.check(_lib.lib().f2_sigma_rdm(self._h, kind, left.data_ptr(), right.data_ptr(), out.data_ptr(), stream))
        else:
            _lib.check(_lib.lib().f2_sigma_rdm_part(self._h, kind, left.data_ptr(), right.data_ptr(), out.data_ptr(),
                                                    int(part[0]), int(part[1]), stream))
        return out.view(shape) if device_out else out.cpu().numpy().reshape(shape)

    def a_1rdm(self, C_left, C_right, **kw) -> np.ndarray:
        """<L|a+_p a_q|R> for alpha, (norb, norb) (CISigmaBuilder::compute_a_1rdm, ci_sigma_builder_1rdm.cc:9-65)"""
        n = self.ci_strings.norb
        return self._rdm(0, C_left, C_right, (n, n), **kw)

    def b_1rdm(self, C_left, C_right, **kw) -> np.ndarray:
        n = self.ci_strings.norb
        return self._rdm(1, C_left, C_right, (n, n), **kw)

    def sf_1rdm(self, C_left, C_right) -> np.ndarray:
        """spin-free 1-RDM = alpha + beta (ci_sigma_builder_1rdm.cc:71-81)"""
        return self.a_1rdm(C_left, C_right) + self.b_1rdm(C_left, C_right)

    def _npairs(self):
        n = self.ci_strings.norb
        return n * (n - 1) // 2

    def aa_2rdm(self, C_left, C_right, **kw) -> np.ndarray:
        """<L|a+_p a+_q a_s a_r|R>, p > q, r > s, packed (npairs, npairs) (compute_aa_2rdm,
        ci_sigma_builder_2rdm.cc:45-125); (0, 0) with fewer than two orbitals like the reference"""
        m = self._npairs() if self.ci_strings.norb >= 2 else 0
        return self._rdm(3, C_left, C_right, (m, m), **kw)

    def bb_2rdm(self, C_left, C_right, **kw) -> np.ndarray:
        m = self._npairs() if self.ci_strings.norb >= 2 else 0
        return self._rdm(4, C_left, C_right, (m, m), **kw)

    def ab_2rdm(self, C_left, C_right, **kw) -> np.ndarray:
        """<L|a+_u(alpha) a+_v(beta) a_y(beta) a_x(alpha)|R>, (norb,) * 4 (compute_ab_2rdm,
        ci_sigma_builder_2rdm.cc:135-201)"""
        n = self.ci_strings.norb
        return self._rdm(5, C_left, C_right, (n, n, n, n), **kw)

    def sf_2rdm(self, C_left, C_right) -> np.ndarray:
        """spin-free 2-RDM (compute_sf_2rdm, ci_sigma_builder_2rdm.cc:203-254)"""
        n = self.ci_strings.norb
        ab = self.ab_2rdm(C_left, C_right)
        if n < 2:
            return assemble_sf_2rdm(n, ab, None, None)
        return assemble_sf_2rdm(n, ab, self.aa_2rdm(C_left, C_right), self.bb_2rdm(C_left, C_right))

    def sf_2cumulant(self, C_left, C_right) -> np.ndarray:
        """L2 = G2 - G1 x G1 + 1/2 exchange (compute_sf_2cumulant, ci_sigma_builder_2rdm.cc:256-279)"""
        g1 = self.sf_1rdm(C_left, C_right)
        return (self.sf_2rdm(C_left, C_right) - np.einsum("pr,qs->pqrs", g1, g1)
                + 0.5 * np.einsum("ps,qr->pqrs", g1, g1))

    # ---- three-body blocks, cumulant, transition 1-RDMs (forte2/api/ci_helpers_api.cc:131-145)
    def _ntriplets(self):
        n = self.ci_strings.norb
        return n * (n - 1) * (n - 2) // 6

    def aaa_3rdm(self, C_left, C_right) -> np.ndarray:
        """<L|a+_p a+_q a+_r a_u a_t a_s|R>, p > q > r, s > t > u, packed (ntriplets, ntriplets) (compute_aaa_3rdm,
        ci_sigma_builder_3rdm.cc:84-162); (0, 0) with fewer than three orbitals like the reference"""
        m = self._ntriplets() if self.ci_strings.norb >= 3 else 0
        return self._rdm(8, C_left, C_right, (m, m))

    def bbb_3rdm(self, C_left, C_right) -> np.ndarray:
        m = self._ntriplets() if self.ci_strings.norb >= 3 else 0
        return self._rdm(11, C_left, C_right, (m, m))

    def aab_3rdm(self, C_left, C_right, **kw) -> np.ndarray:
        """<L|a+_p a+_q b+_r b_u a_t a_s|R>, [(p > q), r, (s > t), u] (compute_aab_3rdm,
        ci_sigma_builder_3rdm.cc:172-262)"""
        n = self.ci_strings.norb
        if n < 2:
            return np.zeros((0, 0, 0, 0))
        return self._rdm(9, C_left, C_right, (self._npairs(), n, self._npairs(), n), **kw)

    def abb_3rdm(self, C_left, C_right) -> np.ndarray:
        """<L|a+_p b+_q b+_r b_u b_t a_s|R>, [p, (q > r), s, (t > u)] (compute_abb_3rdm,
        ci_sigma_builder_3rdm.cc:264-318)"""
        n = self.ci_strings.norb
        if n < 2:
            return np.zeros((0, 0, 0, 0))
        return self._rdm(10, C_left, C_right, (n, self._npairs(), n, self._npairs()))

    def sf_3rdm(self, C_left, C_right) -> np.ndarray:
        """spin-free 3-RDM, (norb,) * 6 (compute_sf_3rdm, ci_sigma_builder_3rdm.cc:320-495): the four spin blocks from
        the device, unpacked and summed on the host"""
        n = self.ci_strings.norb
        if n < 2:
            return np.zeros((n,) * 6)
        sss = None
        if n >= 3:
            sss = self.aaa_3rdm(C_left, C_right) + self.bbb_3rdm(C_left, C_right)
        return assemble_sf_3rdm(n, self.aab_3rdm(C_left, C_right), self.abb_3rdm(C_left, C_right), sss)

    def sf_3cumulant(self, C_left, C_right) -> np.ndarray:
        """L3 = G3 - sum G1 G2 + sum G1 G1 G1 with the spin-free prefactors of compute_sf_3cumulant
        (ci_sigma_builder_3rdm.cc:497-540)"""
        g1 = self.sf_1rdm(C_left, C_right)
        g2 = self.sf_2rdm(C_left, C_right)
        return sf_3cumulant_from(g1, g2, self.sf_3rdm(C_left, C_right))

    def _trdm(self, spin, right, C_left, C_right):
        import torch
        if not isinstance(right, CISigmaBuilder):
            raise TypeError("sigmabuilder_right must be a CISigmaBuilder")
        dev = torch.device("cuda", self.device)

        def vec(x, ndet, name):
            t = x if _is_torch_tensor(x) else torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
            if not t.is_cuda or t.dtype != torch.float64 or t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
                raise TypeError(f"{name} must be a one-dimensional contiguous float64 vector")
            if t.numel() != ndet:
                raise RuntimeError("Size of CI vector does not match the number of determinants.")
            return t

        left = vec(C_left, self.ci_strings.ndet, "C_left")
        rvec = vec(C_right, right.ci_strings.ndet, "C_right")
        n = self.ci_strings.norb
        out = torch.zeros(max(n * n, 1), dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(_lib.lib().f2_sigma_trdm1(self._h, right._h, spin, left.data_ptr(), rvec.data_ptr(), out.data_ptr(),
                                            stream))
        return out[: n * n].cpu().numpy().reshape(n, n)

    def a_1trdm(self, sigmabuilder_right, C_left, C_right) -> np.ndarray:
        """<L|a+_p a_q|R> (alpha) between this builder's CI space (left) and another builder's (right): same orbitals and
        electron numbers, different symmetry / GAS restrictions (compute_a_1trdm, ci_sigma_builder_1trdm.cc:11-113)"""
        return self._trdm(0, sigmabuilder_right, C_left, C_right)

    def b_1trdm(self, sigmabuilder_right, C_left, C_right) -> np.ndarray:
        return self._trdm(1, sigmabuilder_right, C_left, C_right)

    def sf_1trdm(self, sigmabuilder_right, C_left, C_right) -> np.ndarray:
        """alpha + beta (compute_sf_1trdm, ci_sigma_builder_1trdm.cc:121-133)"""
        return (self._trdm(0, sigmabuilder_right, C_left, C_right)
                + self._trdm(1, sigmabuilder_right, C_left, C_right))

    def make_sparse_state(self, C, threshold: float = 1e-12):
        """{determinant (alpha word, beta word): coefficient} of the entries with |c| > threshold
        (CISigmaBuilder::make_sparse_state, ci_sigma_builder.cc; the reference returns its SparseState type, which is
        outside this path: a plain dict keyed by the two occupation words stands in)"""
        c = C.detach().cpu().numpy() if _is_torch_tensor(C) else np.asarray(C, dtype=np.float64)
        if c.shape != (self.ci_strings.ndet,):
            raise RuntimeError("Size of CI vector does not match the number of determinants.")
        dets = self.ci_strings.make_determinants()
        keep = np.nonzero(np.abs(c) > threshold)[0]
        return {(int(dets[k][0]), int(dets[k][1])): float(c[k]) for k in keep}

    def _debug_rdm(self, *args, **kwargs):
        raise NotImplementedError("the *_debug RDM builders of the reference (sparse-operator based cross-checks, "
                                  "api/ci_helpers_api.cc:148-190) are not part of the device path")

    a_1rdm_debug = aa_2rdm_debug = ab_2rdm_debug = aaa_3rdm_debug = aab_3rdm_debug = abb_3rdm_debug = _debug_rdm
    aaaa_4rdm_debug = aaab_4rdm_debug = aabb_4rdm_debug = abbb_4rdm_debug = _debug_rdm

    def form_Hdiag_det(self, spin_averaged: bool = False):
        """<I|H|I> for every determinant (forte2/determinant/slater_rules.cc:86-121) as a CUDA tensor.
        spin_averaged: the average over the determinants of each orbital configuration instead (a diagonal that
        commutes with S^2; the preconditioner of determinant-basis solves with spin control)."""
        import torch
        out = torch.empty(self.ci_strings.ndet, dtype=torch.float64, device=f"cuda:{self.device}")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        if spin_averaged:
            _lib.check(_lib.lib().f2_sigma_hdiag_spin_averaged(self._h, out.data_ptr(), 0, 1, stream))
        else:
            _lib.check(_lib.lib().f2_sigma_hdiag(self._h, out.data_ptr(), stream))
        return out


def _slater_block(strings: CIStrings, E, H, V, idx) -> np.ndarray:
    """<d|H|d'> for the determinants at the CI-vector addresses idx (host, f2_slater_matrix:
    SlaterRules::slater_rules, forte2/determinant/slater_rules.cc:123-245)"""
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    block = np.zeros((idx.size, idx.size))
    _lib.check(_lib.lib().f2_slater_matrix(strings._h, float(E), H.ctypes.data_as(_lib.f64p),
                                           V.ctypes.data_as(_lib.f64p), idx.ctypes.data_as(_lib.i64p),
                                           int(idx.size), block.ctypes.data_as(_lib.f64p)))
    return block


def exact_hdiag_csf(strings: CIStrings, E, H, V, spin_adapter) -> np.ndarray:
    """<I|H|I> = sum_{d d'} c_dI c_d'I <d|H|d'> for every CSF I (CISigmaBuilder::form_Hdiag_csf with
    spin_adapt_full_preconditioner = true -> energy_csf, forte2/ci/ci_sigma_builder.cc:326-332): host work,
    O(sum over configurations of ndet_conf^2) Slater-rule matrix elements (f2_slater_matrix; needs no GPU).
    The CSFs of one configuration share their determinants, so the determinant block is formed once per
    configuration."""
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    bounds, det, coeff = spin_adapter.table()
    out = np.zeros(spin_adapter.ncsf)
    for first, last in spin_adapter.conf_to_csfs_range():
        if last <= first:
            continue
        dets = np.unique(det[bounds[first]:bounds[last]])
        block = _slater_block(strings, E, H, V, dets)
        for i in range(first, last):
            where = np.searchsorted(dets, det[bounds[i]:bounds[i + 1]])
            c = coeff[bounds[i]:bounds[i + 1]]
            out[i] = c @ block[np.ix_(where, where)] @ c
    return out


def _ncsf(spin_adapter) -> int:
    return int(spin_adapter.ncsf() if callable(spin_adapter.ncsf) else spin_adapter.ncsf)


def csf_matrix_element(strings: CIStrings, E, H, V, spin_adapter, I: int, J: int) -> float:
    """<I|H|J> between two CSFs = sum_{d d'} c_dI c_d'J <d|H|d'> (ci_sigma_builder.cc:278-315); host"""
    row_I, row_J = spin_adapter.csf(int(I)), spin_adapter.csf(int(J))
    idx = np.array(sorted({int(d) for d, _ in row_I} | {int(d) for d, _ in row_J}), dtype=np.int64)
    pos = {int(d): k for k, d in enumerate(idx)}
    c_I, c_J = np.zeros(idx.size), np.zeros(idx.size)
    for d, c in row_I:
        c_I[pos[int(d)]] += c
    for d, c in row_J:
        c_J[pos[int(d)]] += c
    return float(c_I @ _slater_block(strings, E, H, V, idx) @ c_J)


def csf_hamiltonian(strings: CIStrings, E, H, V, spin_adapter) -> np.ndarray:
    """T^t H_det T over the determinants the CSFs use (ci_sigma_builder.cc:372-392); host"""
    ncsf = _ncsf(spin_adapter)
    rows = [spin_adapter.csf(i) for i in range(ncsf)]
    idx = np.array(sorted({int(d) for row in rows for d, _ in row}), dtype=np.int64)
    pos = {int(d): k for k, d in enumerate(idx)}
    T = np.zeros((idx.size, ncsf))
    for i, row in enumerate(rows):
        for d, c in row:
            T[pos[int(d)], i] += c
    Hcsf = T.T @ _slater_block(strings, E, H, V, idx) @ T
    return 0.5 * (Hcsf + Hcsf.T)


def _pair_expansion(n):
    """E[pq, p, q] = +1, E[pq, q, p] = -1 for p > q, pq = p (p - 1) / 2 + q"""
    m = n * (n - 1) // 2
    E = np.zeros((m, n, n))
    p, q = np.tril_indices(n, -1)
    k = p * (p - 1) // 2 + q
    E[k, p, q] = 1.0
    E[k, q, p] = -1.0
    return E


def _triplet_expansion(n):
    """E[pqr, a, b, c] = sign of the permutation (a, b, c) of p > q > r, pqr = p (p-1) (p-2) / 6 + q (q-1) / 2 + r"""
    from itertools import combinations
    m = n * (n - 1) * (n - 2) // 6
    E = np.zeros((m, n, n, n))
    for r, q, p in combinations(range(n), 3):
        k = p * (p - 1) * (p - 2) // 6 + q * (q - 1) // 2 + r
        E[k, p, q, r] = E[k, r, p, q] = E[k, q, r, p] = 1.0
        E[k, p, r, q] = E[k, r, q, p] = E[k, q, p, r] = -1.0
    return E


def assemble_sf_3rdm(n, aab, abb, sss) -> np.ndarray:
    """spin-free 3-RDM from the packed spin blocks (the scatter loops of compute_sf_3rdm,
    ci_sigma_builder_3rdm.cc:337-493): aab [(p>q), r, (s>t), u], abb [p, (q>r), s, (t>u)], sss = aaa + bbb
    [(p>q>r), (s>t>u)] (None with fewer than three orbitals)"""
    P = _pair_expansion(n)
    m = P.shape[0]
    # X[p,q,r,s,t,u] = <a+_p a+_q b+_r b_u a_t a_s>, antisymmetric in (p,q) and (s,t)
    X = np.tensordot(P.reshape(m, n * n).T, aab.reshape(m, n * m * n), axes=1).reshape(n, n, n, m, n)
    X = np.tensordot(X, P, axes=([3], [0])).transpose(0, 1, 2, 4, 5, 3)  # p q r u s t -> p q r s t u
    # Y[p,q,r,s,t,u] = <a+_p b+_q b+_r b_u b_t a_s>, antisymmetric in (q,r) and (t,u)
    Y = np.tensordot(abb.reshape(n, m, n, m), P, axes=([3], [0]))  # p, qr, s, t, u
    Y = np.tensordot(P, Y, axes=([0], [1])).transpose(2, 0, 1, 3, 4, 5)  # q r p s t u -> p q r s t u
    out = (X + X.transpose(0, 2, 1, 3, 5, 4) + X.transpose(2, 0, 1, 5, 3, 4)
           + Y + Y.transpose(1, 0, 2, 4, 3, 5) + Y.transpose(1, 2, 0, 4, 5, 3))
    if sss is not None and n >= 3:
        T = _triplet_expansion(n)
        k = T.shape[0]
        Tm = T.reshape(k, n ** 3)
        out = out + (Tm.T @ sss @ Tm).reshape((n,) * 6)
    return np.ascontiguousarray(out)


def sf_3cumulant_from(g1, g2, g3) -> np.ndarray:
    """compute_sf_3cumulant (ci_sigma_builder_3rdm.cc:497-540) from the spin-free 1-, 2- and 3-RDM"""
    e = np.einsum
    return (g3 - e("ps,qrtu->pqrstu", g1, g2) - e("qt,prsu->pqrstu", g1, g2) - e("ru,pqst->pqrstu", g1, g2)
            + 0.5 * e("pt,qrsu->pqrstu", g1, g2) + 0.5 * e("pu,qrts->pqrstu", g1, g2)
            + 0.5 * e("qs,prtu->pqrstu", g1, g2) + 0.5 * e("qu,prst->pqrstu", g1, g2)
            + 0.5 * e("rs,pqut->pqrstu", g1, g2) + 0.5 * e("rt,pqsu->pqrstu", g1, g2)
            + 2.0 * e("ps,qt,ru->pqrstu", g1, g1, g1) - e("ps,qu,rt->pqrstu", g1, g1, g1)
            - e("pu,qt,rs->pqrstu", g1, g1, g1) - e("pt,qs,ru->pqrstu", g1, g1, g1)
            + 0.5 * e("pt,qu,rs->pqrstu", g1, g1, g1) + 0.5 * e("pu,qs,rt->pqrstu", g1, g1, g1))


def assemble_sf_2rdm(norb: int, ab: np.ndarray, aa, bb) -> np.ndarray:
    """G[p,q,r,s] = ab[p,q,r,s] + ab[q,p,s,r] + the antisymmetric unpacking of the packed same-spin blocks
    (ci_sigma_builder_2rdm.cc:215-252); aa / bb are (npairs, npairs) over p > q, r > s or None"""
    out = ab + ab.transpose(1, 0, 3, 2)
    if norb < 2:
        return out
    p, q = np.tril_indices(norb, k=-1)  # p > q in the order of pair_index_gt: p (p - 1) / 2 + q
    for ss in (aa, bb):
        if ss is None:
            continue
        P, R = p[:, None], p[None, :]
        Q, S = q[:, None], q[None, :]
        out[P, Q, R, S] += ss
        out[Q, P, R, S] -= ss
        out[P, Q, S, R] -= ss
        out[Q, P, S, R] += ss
    return out
