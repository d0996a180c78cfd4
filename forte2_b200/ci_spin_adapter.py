"""CISpinAdapter: the Python face of forte2.lib.ci_helpers.CISpinAdapter
(forte2/api/ci_helpers_api.cc:190-201, forte2/determinant/ci_spin_adapter.h:48) on one B200.

Same constructor (`twoS, twoMs, norb`), `prepare_couplings`, `csf_C_to_det_C`, `det_C_to_csf_C`,
`nconf`, `ncsf`, `set_log_level`.  `prepare_couplings` takes the `CIStrings` object instead of the
O(ndet) host list of `Determinant` objects the reference needs (the addresses come from the string
tables); passing `strings.make_determinants()` of the same space is accepted for source compatibility.

The transforms accept torch CUDA float64 tensors (zero copy, asynchronous on the current stream) or
numpy vectors as the reference does (staged through the device inside the C ABI, synchronous).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .ci_strings import CIStrings


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch")


class CISpinAdapter:
    def __init__(self, twoS: int, twoMs: int, norb: int):
        self.twoS, self.twoMs, self.norb = int(twoS), int(twoMs), int(norb)
        self._h = None
        self.ci_strings = None
        self.log_level = 3

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.f2_adapter_destroy(h)

    def set_log_level(self, level: int):
        self.log_level = int(level)

    def prepare_couplings(self, dets, strings: CIStrings | None = None):
        """`dets`: a forte2_b200.CIStrings (preferred), or the determinant array of one passed as `strings`."""
        if isinstance(dets, CIStrings):
            strings = dets
        elif strings is None:
            raise TypeError("prepare_couplings needs the CIStrings of the determinant space (pass it instead of, "
                            "or together with, the determinant list)")
        elif len(dets) != strings.ndet:
            raise ValueError("the determinant list does not belong to the given CIStrings")
        if strings.norb != self.norb:
            raise ValueError(f"the string space has {strings.norb} orbitals, the adapter {self.norb}")
        if self._h:
            _lib.lib().f2_adapter_destroy(self._h)
            self._h = None
        self._h = _lib.lib().f2_adapter_create(strings._h, self.twoS, self.twoMs)
        if not self._h:
            _lib.check(_lib.lib().f2_last_status() or 2)
        self.ci_strings = strings
        info = np.zeros(4, dtype=np.int64)
        _lib.check(_lib.lib().f2_adapter_info(self._h, info.ctypes.data_as(_lib.i64p)))
        self._nconf, self._ncsf, self._ndet, self._ncoupling = (int(x) for x in info)

    def _need(self):
        if not self._h:
            raise RuntimeError("CISpinAdapter: prepare_couplings has not been called")

    @property
    def nconf(self) -> int:
        self._need()
        return self._nconf

    @property
    def ncsf(self) -> int:
        self._need()
        return self._ncsf

    @property
    def ndet(self) -> int:
        self._need()
        return self._ndet

    @property
    def ncoupling(self) -> int:
        self._need()
        return self._ncoupling

    def table(self):
        """(bounds, det, coeff): CSF i = sum over [bounds[i], bounds[i+1]) of coeff * Determinant(det)."""
        self._need()
        bounds = np.zeros(self._ncsf + 1, dtype=np.int64)
        det = np.zeros(max(self._ncoupling, 1), dtype=np.int64)
        coeff = np.zeros(max(self._ncoupling, 1))
        _lib.lib().f2_adapter_table(self._h, bounds.ctypes.data_as(_lib.i64p), det.ctypes.data_as(_lib.i64p),
                                    coeff.ctypes.data_as(_lib.f64p), det.size)
        return bounds, det[: self._ncoupling], coeff[: self._ncoupling]

    def conf_to_csfs_range(self):
        self._need()
        first = np.zeros(self._nconf + 1, dtype=np.int64)
        _lib.check(_lib.lib().f2_adapter_conf_ranges(self._h, first.ctypes.data_as(_lib.i64p)))
        return [(int(a), int(b)) for a, b in zip(first[:-1], first[1:])]

    def _transform(self, src, dst, n_src, n_dst, dev_fn, host_fn, name):
        self._need()
        if isinstance(src, np.ndarray) and isinstance(dst, np.ndarray):
            for v in (src, dst):
                if v.dtype != np.float64 or v.ndim != 1:
                    raise TypeError("vectors must be one-dimensional float64")
            # strided views are accepted like in the reference, which reads and writes through .view()
            # (ci_spin_adapter.cc:69-97): forte2's Davidson closure passes columns of its C-ordered basis,
            # Bblock[:, i] and Sblock[:, i] (forte2/ci/ci.py:231-233)
            s = src if src.size <= 1 or src.strides[0] == 8 else np.ascontiguousarray(src)
            d = dst if dst.size <= 1 or dst.strides[0] == 8 else np.empty(dst.shape[0])
            _lib.check(host_fn(self._h, s.ctypes.data_as(_lib.f64p), s.shape[0], d.ctypes.data_as(_lib.f64p),
                               d.shape[0]))
            if d is not dst:
                dst[:] = d
            return
        import torch
        if not _is_torch_tensor(src):
            src = torch.from_dlpack(src)
        if not _is_torch_tensor(dst):
            dst = torch.from_dlpack(dst)
        for v, n in ((src, n_src), (dst, n_dst)):
            if not v.is_cuda or v.dtype != torch.float64 or v.dim() != 1:
                raise TypeError(f"{name}: vectors must be one-dimensional float64 CUDA tensors")
            if v.numel() != n:
                raise ValueError(f"{name}: expected a vector of {n} elements, got {v.numel()}")
            if v.numel() > 1 and v.stride(0) != 1:
                raise RuntimeError("np_vector must have stride 1 to convert to std::span.")
        if src.device != dst.device:
            raise ValueError(f"{name}: vectors live on different devices")
        stream = torch.cuda.current_stream(src.device).cuda_stream
        _lib.check(dev_fn(self._h, src.data_ptr(), dst.data_ptr(), src.device.index, stream))

    def csf_C_to_det_C(self, csf_C, det_C) -> None:
        L = _lib.lib()
        self._transform(csf_C, det_C, self._ncsf if self._h else 0, self._ndet if self._h else 0,
                        L.f2_adapter_csf_to_det,
                        lambda h, s, ns, d, nd: L.f2_adapter_csf_to_det_host(h, s, ns, d, nd), "csf_C_to_det_C")

    def det_C_to_csf_C(self, det_C, csf_C) -> None:
        L = _lib.lib()
        self._transform(det_C, csf_C, self._ndet if self._h else 0, self._ncsf if self._h else 0,
                        L.f2_adapter_det_to_csf,
                        lambda h, s, ns, d, nd: L.f2_adapter_det_to_csf_host(h, s, ns, d, nd), "det_C_to_csf_C")

    def csf(self, i: int):
        """[(determinant address, coefficient)] of CSF i (CISpinAdapter::csf, ci_spin_adapter.h:168)"""
        bounds, det, coeff = self._cached_table()
        if not 0 <= i < self._ncsf:
            raise IndexError(i)
        return list(zip(det[bounds[i]:bounds[i + 1]].tolist(), coeff[bounds[i]:bounds[i + 1]].tolist()))

    def _cached_table(self):
        if getattr(self, "_table", None) is None or self._table_of is not self._h:
            self._table, self._table_of = self.table(), self._h
        return self._table
