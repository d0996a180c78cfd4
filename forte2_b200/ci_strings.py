"""CIStrings: the Python face of forte2.lib.ci_helpers.CIStrings (forte2/api/ci_helpers_api.cc:52-85)
backed by the C ABI.  Same constructor signature, same read-only properties, same exceptions."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

KINDS = {
    "strings": 0, "strings_1h": 1, "strings_2h": 2, "strings_3h": 3, "strings_1h1p": 4,
    "vo": 10, "vo2": 11, "1h": 12, "1h2": 13, "2h": 14, "3h": 15,
}
# device-format tables as the kernels read them (csrc/strings.hpp DeviceLists), for host models of kernels in tests/
DEV_KINDS = {
    "dev_as_ptr": 20, "dev_as_ent": 21, "dev_bs_ent": 22, "dev_h2_ptr": 23, "dev_h2_ent": 24, "dev_cls": 25,
    "dev_loc": 26, "dev_blk_off": 27, "dev_ncls": 28,
    # hole lists of the RDM path (csrc/strings.hpp HoleLists), level 0..3: CSR pointer and packed entries
    "hole_ptr0": 30, "hole_ptr1": 31, "hole_ptr2": 32, "hole_ptr3": 33,
    "hole_ent0": 34, "hole_ent1": 35, "hole_ent2": 36, "hole_ent3": 37,
}
ROW_WIDTH = {"vo": 7, "vo2": 6, "1h": 6, "1h2": 6, "2h": 7, "3h": 8}


def _i32(a):
    return a.ctypes.data_as(_lib.i32p)


def _i64(a):
    return a.ctypes.data_as(_lib.i64p)


class CIStrings:
    def __init__(self, na, nb, symmetry, orbital_symmetry, gas_min, gas_max):
        if not isinstance(na, (int, np.integer)) or not isinstance(nb, (int, np.integer)):
            raise TypeError("na and nb must be integers")
        if na < 0:
            raise ValueError("Number of alpha electrons must be non-negative.")
        if nb < 0:
            raise ValueError("Number of beta electrons must be non-negative.")
        if not isinstance(orbital_symmetry, (list, tuple)) or any(
                not isinstance(s, (list, tuple, np.ndarray)) for s in orbital_symmetry):
            raise TypeError("orbital_symmetry must be a list of lists of irrep labels (one per GAS space)")
        L = _lib.lib()
        sizes = np.array([len(s) for s in orbital_symmetry], dtype=np.int32).reshape(-1)
        flat = np.array([int(x) for s in orbital_symmetry for x in s], dtype=np.int32).reshape(-1)
        gmin = np.array(list(gas_min), dtype=np.int32).reshape(-1)
        gmax = np.array(list(gas_max), dtype=np.int32).reshape(-1)
        self._h = L.f2_strings_create(int(na), int(nb), int(symmetry), len(sizes), _i32(sizes), _i32(flat),
                                      len(gmin), _i32(gmin), len(gmax), _i32(gmax))
        if not self._h:
            msg = _lib.last_error()
            raise (IndexError if "outside the irrep range" in msg else ValueError)(msg)
        info = np.zeros(12, dtype=np.int64)
        L.f2_strings_info(self._h, _i64(info))
        (self.na, self.nb, self.nas, self.nbs, self.ndet, self.norb, self.nirrep, self.ngas_spaces,
         self.nblocks, self.symmetry, self.nclasses_a, self.nclasses_b) = (int(x) for x in info)
        self.gas_size = [int(x) for x in sizes]
        self._orbital_symmetry = [list(map(int, s)) for s in orbital_symmetry]

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and _lib._lib is not None:
            _lib._lib.f2_strings_destroy(h)

    # ---- properties of the reference binding ----
    def _occ(self, which):
        L = _lib.lib()
        n = L.f2_strings_occupations(self._h, which, None, 0)
        out = np.zeros(n, dtype=np.int64)
        L.f2_strings_occupations(self._h, which, _i64(out), n)
        return out

    @property
    def gas_alpha_occupations(self):
        return [[int(x) for x in row] for row in self._occ(0).reshape(-1, 6)]

    @property
    def gas_beta_occupations(self):
        return [[int(x) for x in row] for row in self._occ(1).reshape(-1, 6)]

    @property
    def gas_occupations(self):
        return [(int(a), int(b)) for a, b in self._occ(2).reshape(-1, 2)]

    def determinant(self, address):
        out = np.zeros(2, dtype=np.uint64)
        _lib.check(_lib.lib().f2_strings_determinant(self._h, int(address), out.ctypes.data_as(_lib.u64p)))
        return int(out[0]), int(out[1])

    def determinant_index(self, d):
        a, b = d
        r = _lib.lib().f2_strings_determinant_index(self._h, int(a), int(b))
        if r < 0:
            raise IndexError(_lib.last_error())
        return int(r)

    def make_determinants(self):
        """(ndet, 2) uint64: alpha and beta occupation words in CI-vector order."""
        out = np.zeros((self.ndet, 2), dtype=np.uint64)
        _lib.lib().f2_strings_determinants(self._h, out.ctypes.data_as(_lib.u64p))
        return out

    # ---- additions used by the tests and the multi-GPU driver ----
    def blocks(self):
        """rows (class_Ia, class_Ib, offset, nIa, nIb)"""
        out = np.zeros((self.nblocks, 5), dtype=np.int64)
        _lib.lib().f2_strings_blocks(self._h, _i64(out))
        return out

    def occupations(self, which):
        return self._occ(which).reshape(-1, 6 if which < 2 else 2)

    def build_lists(self, device=None):
        """build the replacement lists now: on GPU `device` (CUDA kernels, csrc/lists_device.cu) or, with None, on the
        host; CISigmaBuilder does the former on its own when the lists do not exist yet"""
        _lib.check(_lib.lib().f2_strings_build_lists(self._h, -1 if device is None else int(device)))

    def export(self, kind, spin):
        L = _lib.lib()
        k = KINDS[kind] if kind in KINDS else DEV_KINDS[kind]
        n = L.f2_strings_export(self._h, k, spin, None, 0)
        if n < 0:
            raise RuntimeError(_lib.last_error())
        out = np.zeros(n, dtype=np.int64)
        L.f2_strings_export(self._h, k, spin, _i64(out), n)
        return out.reshape(-1, ROW_WIDTH[kind]) if kind in ROW_WIDTH else out
