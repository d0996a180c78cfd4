"""ctypes loader for oracle/_ref/libforte2_ref.so — TEST INFRASTRUCTURE, not the product.

The library is the reference's own C++ CI code (forte2/ci/*.cc, forte2/determinant/*.cc) compiled
in place by oracle/Makefile together with the C-ABI glue in oracle/ref_cabi.cc.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "_ref" / "libforte2_ref.so"

KINDS = {
    "strings": 0, "strings_1h": 1, "strings_2h": 2, "strings_3h": 3, "strings_1h1p": 4,
    "vo": 10, "vo2": 11, "1h": 12, "1h2": 13, "2h": 14, "3h": 15,
}
ROW_WIDTH = {"vo": 7, "vo2": 6, "1h": 6, "1h2": 6, "2h": 7, "3h": 8}

_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built: run `make -C oracle ref` where /root/reference exists")
        # the OpenBLAS inside the opencv wheel needs its private libgfortran/libquadmath, which sit
        # next to it without an rpath: preload them so the dynamic loader finds them by soname
        import sysconfig
        blasdir = Path(sysconfig.get_paths()["purelib"]) / "opencv_python_headless.libs"
        for pat in ("libquadmath-*.so*", "libgfortran-*.so*", "libopenblasp-*.so*"):
            for p in sorted(blasdir.glob(pat)):
                C.CDLL(str(p), mode=C.RTLD_GLOBAL)
        L = C.CDLL(str(LIB_PATH))
        vp, i, i64p, dp, i64 = C.c_void_p, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.c_int64
        ip = C.POINTER(C.c_int)
        L.ref_last_error.restype = C.c_char_p
        L.ref_strings_create.restype = vp
        L.ref_strings_create.argtypes = [i, i, i, i, ip, ip, i, ip, i, ip]
        L.ref_strings_destroy.argtypes = [vp]
        L.ref_strings_info.argtypes = [vp, i64p]
        L.ref_strings_blocks.argtypes = [vp, i64p]
        L.ref_strings_occupations.restype = i64
        L.ref_strings_occupations.argtypes = [vp, i, i64p, i64]
        L.ref_strings_export.restype = i64
        L.ref_strings_export.argtypes = [vp, i, i, i64p, i64]
        L.ref_strings_determinants.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.ref_sigma_create.restype = vp
        L.ref_sigma_create.argtypes = [vp, C.c_double, dp, dp, i]
        L.ref_sigma_destroy.argtypes = [vp]
        L.ref_sigma_set_algorithm.argtypes = [vp, C.c_char_p]
        L.ref_sigma_get_algorithm.restype = C.c_char_p
        L.ref_sigma_get_algorithm.argtypes = [vp]
        L.ref_sigma_set_memory.argtypes = [vp, i]
        L.ref_sigma_apply.argtypes = [vp, dp, dp, i64]
        L.ref_sigma_times.argtypes = [vp, dp]
        L.ref_det_energies.argtypes = [vp, dp]
        L.ref_dense_hamiltonian.argtypes = [vp, C.c_double, dp]
        L.ref_adapter_create.restype = vp
        L.ref_adapter_create.argtypes = [vp, i, i]
        L.ref_adapter_destroy.argtypes = [vp]
        L.ref_adapter_ncsf.restype = i64
        L.ref_adapter_ncsf.argtypes = [vp]
        L.ref_adapter_ndet.restype = i64
        L.ref_adapter_ndet.argtypes = [vp]
        L.ref_adapter_csf_to_det.argtypes = [vp, dp, i64, dp, i64]
        L.ref_adapter_det_to_csf.argtypes = [vp, dp, i64, dp, i64]
        L.ref_adapter_table.restype = i64
        L.ref_adapter_table.argtypes = [vp, i64p, i64p, dp, i64]
        L.ref_hdiag_csf.argtypes = [vp, vp, i, dp]
        L.ref_rdm.restype = i64
        L.ref_rdm.argtypes = [vp, i, dp, dp, i64, dp, i64]
        L.ref_trdm.restype = i64
        L.ref_trdm.argtypes = [vp, vp, i, dp, i64, dp, i64, dp, i64]
        L.ref_form_H_csf.restype = i
        L.ref_form_H_csf.argtypes = [vp, vp, dp]
        L.ref_slater_rules_csf.restype = C.c_double
        L.ref_slater_rules_csf.argtypes = [vp, vp, i64, i64]
        L.ref_set_log_level.argtypes = [i]
        L.ref_set_log_level(0)
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def _i64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


_EXC = {1: ValueError, 2: RuntimeError, 3: OverflowError}


def _check(status):
    if status:
        raise _EXC.get(status, RuntimeError)(lib().ref_last_error().decode())


class RefCIStrings:
    """forte2.lib.ci_helpers.CIStrings as compiled from the reference (api/ci_helpers_api.cc:52-85)."""

    def __init__(self, na, nb, symmetry, orbital_symmetry, gas_min, gas_max):
        L = lib()
        sizes = np.array([len(s) for s in orbital_symmetry], dtype=np.int32)
        flat = np.array([x for s in orbital_symmetry for x in s], dtype=np.int32).reshape(-1)
        gmin = np.array(gas_min, dtype=np.int32).reshape(-1)
        gmax = np.array(gas_max, dtype=np.int32).reshape(-1)
        self._h = L.ref_strings_create(na, nb, symmetry, len(sizes), _ip(sizes), _ip(flat),
                                       len(gmin), _ip(gmin), len(gmax), _ip(gmax))
        if not self._h:
            raise ValueError(L.ref_last_error().decode())
        info = np.zeros(10, dtype=np.int64)
        L.ref_strings_info(self._h, _i64p(info))
        (self.na, self.nb, self.nas, self.nbs, self.ndet, self.norb, self.nirrep, self.ngas_spaces,
         self.nblocks, self.symmetry) = (int(x) for x in info)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_strings_destroy(self._h)
            self._h = None

    def blocks(self):
        out = np.zeros((self.nblocks, 5), dtype=np.int64)
        lib().ref_strings_blocks(self._h, _i64p(out))
        return out

    def occupations(self, which):
        n = lib().ref_strings_occupations(self._h, which, None, 0)
        out = np.zeros(n, dtype=np.int64)
        lib().ref_strings_occupations(self._h, which, _i64p(out), n)
        return out.reshape(-1, 6 if which < 2 else 2)

    def export(self, kind, spin):
        k = KINDS[kind]
        n = lib().ref_strings_export(self._h, k, spin, None, 0)
        if n < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        out = np.zeros(n, dtype=np.int64)
        lib().ref_strings_export(self._h, k, spin, _i64p(out), n)
        if kind in ROW_WIDTH:
            return out.reshape(-1, ROW_WIDTH[kind])
        return out

    def determinants(self):
        out = np.zeros((self.ndet, 2), dtype=np.uint64)
        lib().ref_strings_determinants(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return out


class RefCISigmaBuilder:
    """forte2.lib.ci_helpers.CISigmaBuilder (api/ci_helpers_api.cc:88-109), hot-path subset."""

    def __init__(self, strings: RefCIStrings, E, H, V, log_level=0):
        self.strings = strings
        H = np.ascontiguousarray(H, dtype=np.float64)
        V = np.ascontiguousarray(V, dtype=np.float64)
        self.E = float(E)
        self._h = lib().ref_sigma_create(strings._h, float(E), _dp(H), _dp(V), log_level)
        if not self._h:
            raise RuntimeError(lib().ref_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_sigma_destroy(self._h)
            self._h = None

    def set_algorithm(self, name):
        _check(lib().ref_sigma_set_algorithm(self._h, name.encode()))

    def get_algorithm(self):
        return lib().ref_sigma_get_algorithm(self._h).decode()

    RDM_KINDS = {"a_1rdm": 0, "b_1rdm": 1, "sf_1rdm": 2, "aa_2rdm": 3, "bb_2rdm": 4, "ab_2rdm": 5, "sf_2rdm": 6,
                 "sf_2cumulant": 7, "aaa_3rdm": 8, "aab_3rdm": 9, "abb_3rdm": 10, "bbb_3rdm": 11, "sf_3rdm": 12,
                 "sf_3cumulant": 13}

    def rdm(self, kind, c_left, c_right):
        """CISigmaBuilder::compute_<kind> (ci_sigma_builder_1rdm.cc, ci_sigma_builder_2rdm.cc), shaped as the
        reference returns it: (n, n), (npairs, npairs) for aa/bb (p > q), (n, n, n, n); the 3-RDM blocks
        (ci_sigma_builder_3rdm.cc) as (ntriplets, ntriplets), (npairs, n, npairs, n), (n, npairs, n, npairs), (n,) * 6."""
        k = self.RDM_KINDS[kind]
        cl = np.ascontiguousarray(c_left, dtype=np.float64)
        cr = np.ascontiguousarray(c_right, dtype=np.float64)
        n = lib().ref_rdm(self._h, k, _dp(cl), _dp(cr), cl.shape[0], None, 0)
        if n < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        out = np.zeros(max(n, 1))
        lib().ref_rdm(self._h, k, _dp(cl), _dp(cr), cl.shape[0], _dp(out), n)
        out = out[:n]
        norb = self.strings.norb
        if k <= 2:
            return out.reshape(norb, norb)
        if k <= 4 or k in (8, 11):
            m = int(round(np.sqrt(n)))
            return out.reshape(m, m)
        npairs = norb * (norb - 1) // 2
        if k == 9:
            return out.reshape((npairs, norb, npairs, norb) if n else (0, 0, 0, 0))
        if k == 10:
            return out.reshape((norb, npairs, norb, npairs) if n else (0, 0, 0, 0))
        if k >= 12:
            return out.reshape((norb,) * 6)
        return out.reshape(norb, norb, norb, norb)

    def trdm(self, kind, right_builder, c_left, c_right):
        """CISigmaBuilder::compute_{a,b,sf}_1trdm(sigmabuilder_right, C_left, C_right) (ci_sigma_builder_1trdm.cc:11-133);
        self is the left wave function's builder"""
        k = {"a_1trdm": 0, "b_1trdm": 1, "sf_1trdm": 2}[kind]
        cl = np.ascontiguousarray(c_left, dtype=np.float64)
        cr = np.ascontiguousarray(c_right, dtype=np.float64)
        norb = self.strings.norb
        out = np.zeros(max(norb * norb, 1))
        n = lib().ref_trdm(self._h, right_builder._h, k, _dp(cl), cl.shape[0], _dp(cr), cr.shape[0], _dp(out), out.size)
        if n < 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return out[:n].reshape(norb, norb)

    def set_memory(self, mb):
        _check(lib().ref_sigma_set_memory(self._h, int(mb)))

    def Hamiltonian(self, basis, sigma):
        assert basis.dtype == np.float64 and sigma.dtype == np.float64
        if basis.strides[0] != 8 or sigma.strides[0] != 8:
            raise RuntimeError("np_vector must have stride 1 to convert to std::span.")
        _check(lib().ref_sigma_apply(self._h, _dp(basis), _dp(sigma), basis.shape[0]))

    def avg_build_time(self):
        out = np.zeros(4)
        lib().ref_sigma_times(self._h, _dp(out))
        return list(out)

    def det_energies(self):
        out = np.zeros(self.strings.ndet)
        lib().ref_det_energies(self._h, _dp(out))
        return out + self.E

    def dense_hamiltonian(self):
        n = self.strings.ndet
        out = np.zeros((n, n))
        lib().ref_dense_hamiltonian(self._h, self.E, _dp(out))
        return out


class RefCISpinAdapter:
    """forte2.lib.ci_helpers.CISpinAdapter bound to the determinants of a RefCIStrings."""

    def __init__(self, strings: RefCIStrings, twoS, twoMs):
        self._h = lib().ref_adapter_create(strings._h, twoS, twoMs)
        if not self._h:
            raise RuntimeError(lib().ref_last_error().decode())
        self.ncsf = int(lib().ref_adapter_ncsf(self._h))
        self.ndet = int(lib().ref_adapter_ndet(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_adapter_destroy(self._h)
            self._h = None

    def csf_C_to_det_C(self, csf, det):
        lib().ref_adapter_csf_to_det(self._h, _dp(csf), csf.shape[0], _dp(det), det.shape[0])

    def det_C_to_csf_C(self, det, csf):
        lib().ref_adapter_det_to_csf(self._h, _dp(det), det.shape[0], _dp(csf), csf.shape[0])

    def table(self):
        bounds = np.zeros(self.ncsf + 1, dtype=np.int64)
        n = lib().ref_adapter_table(self._h, _i64p(bounds), None, None, 0)
        idx = np.zeros(n, dtype=np.int64)
        coeff = np.zeros(n)
        lib().ref_adapter_table(self._h, _i64p(bounds), _i64p(idx), _dp(coeff), n)
        return bounds, idx, coeff

    def hdiag_csf(self, builder: RefCISigmaBuilder, full=False):
        out = np.zeros(self.ncsf)
        lib().ref_hdiag_csf(builder._h, self._h, int(full), _dp(out))
        return out

    def form_H_csf(self, builder: RefCISigmaBuilder):
        """CISigmaBuilder::form_H_csf (ci_sigma_builder.cc:372-392); RuntimeError as the reference throws"""
        out = np.zeros((self.ncsf, self.ncsf))
        if lib().ref_form_H_csf(builder._h, self._h, _dp(out)) != 0:
            raise RuntimeError(lib().ref_last_error().decode())
        return out

    def slater_rules_csf(self, builder: RefCISigmaBuilder, I, J):
        """CISigmaBuilder::slater_rules_csf (ci_sigma_builder.cc:278-297)"""
        return float(lib().ref_slater_rules_csf(builder._h, self._h, int(I), int(J)))


def set_threads(n: int | None):
    """The reference reads FORTE_NUM_THREADS_OVERRIDE on every parallel_for (helpers/parallel.h:116-132)."""
    if n is None:
        os.environ.pop("FORTE_NUM_THREADS_OVERRIDE", None)
    else:
        os.environ["FORTE_NUM_THREADS_OVERRIDE"] = str(int(n))
