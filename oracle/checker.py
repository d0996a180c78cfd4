"""TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT.

One entry point for the tests, smoke() and bench.py's cpu_baseline leg: sigma = H C on the CPU from
  * oracle/_ref  — the reference's own C++ builder compiled in place from /root/reference (kind
                   "reference"), when the prebuilt library is present, else
  * oracle/sigma_port.c + oracle/strings_port.py — the restatement (kind "port").
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import ref, strings_port

_HERE = Path(__file__).resolve().parent
PORT_PATH = _HERE / "liboracle_port.so"
_port = None


def port_lib():
    global _port
    if _port is None:
        if not PORT_PATH.exists():
            raise RuntimeError(f"{PORT_PATH} not built: run `make -C oracle port`")
        L = C.CDLL(str(PORT_PATH))
        dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint64)
        L.port_sigma.restype = C.c_int
        L.port_sigma.argtypes = [C.c_int64, up, C.c_int, C.c_double, dp, dp, dp, dp]
        L.port_sigma_parts.restype = C.c_int
        L.port_sigma_parts.argtypes = [C.c_int64, up, C.c_int, C.c_double, dp, dp, dp, dp, C.c_int]
        L.port_energies.argtypes = [C.c_int64, up, C.c_int, C.c_double, dp, dp, dp]
        _port = L
    return _port


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


PART_SCALAR, PART_ALPHA, PART_BETA, PART_ALPHA_BETA, PART_ALL = 1, 2, 4, 8, 15


def port_sigma(dets, E, H, V, c, parts=PART_ALL):
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    out = np.zeros_like(c)
    rc = port_lib().port_sigma_parts(len(c), dets.ctypes.data_as(C.POINTER(C.c_uint64)), H.shape[0], float(E),
                                     _dp(H), _dp(V), _dp(c), _dp(out), int(parts))
    assert rc == 0
    return out


def port_energies(dets, E, H, V):
    dets = np.ascontiguousarray(dets, dtype=np.uint64)
    H = np.ascontiguousarray(H, dtype=np.float64)
    V = np.ascontiguousarray(V, dtype=np.float64)
    out = np.zeros(len(dets))
    port_lib().port_energies(len(dets), dets.ctypes.data_as(C.POINTER(C.c_uint64)), H.shape[0], float(E), _dp(H),
                             _dp(V), _dp(out))
    return out


def sigma(space, E, H, V, c, dets=None, prefer="reference"):
    """Return (sigma, kind).  `dets` is only used by the port when the space is too large for the
    pure-Python enumeration (> 200k determinants)."""
    c = np.ascontiguousarray(c, dtype=np.float64)
    if prefer == "reference" and ref.available():
        s = ref.RefCIStrings(*space)
        b = ref.RefCISigmaBuilder(s, E, H, V)
        b.set_algorithm("hz")
        b.set_memory(4096)
        out = np.zeros_like(c)
        b.Hamiltonian(c, out)
        return out, "reference"
    if dets is None or len(c) <= 200000:
        dets = strings_port.determinants(*space)
    return port_sigma(dets, E, H, V, c), "port"
