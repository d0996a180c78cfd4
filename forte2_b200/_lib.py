"""ctypes binding of forte2_b200/libforte2_b200.so (the C ABI declared in include/forte2_b200.h).

There is deliberately no fallback: if the CUDA library has not been built the import of the
product classes fails with a clear message (python __graft_entry__.py build, or make -C
forte2_b200/csrc).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libforte2_b200.so"
_lib = None

_EXC = {1: ValueError, 2: RuntimeError, 3: OverflowError, 4: IndexError, 5: RuntimeError}

i32p = C.POINTER(C.c_int)
i64p = C.POINTER(C.c_int64)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)

# name -> (restype, argtypes); must list every symbol of include/forte2_b200.h
SIGNATURES = {
    "f2_last_error": (C.c_char_p, []),
    "f2_last_status": (C.c_int, []),
    "f2_abi_version": (C.c_int, []),
    "f2_strings_create": (C.c_void_p, [C.c_int64, C.c_int64, C.c_int, C.c_int, i32p, i32p, C.c_int, i32p,
                                       C.c_int, i32p]),
    "f2_strings_destroy": (None, [C.c_void_p]),
    "f2_strings_info": (C.c_int, [C.c_void_p, i64p]),
    "f2_strings_blocks": (C.c_int, [C.c_void_p, i64p]),
    "f2_strings_occupations": (C.c_int64, [C.c_void_p, C.c_int, i64p, C.c_int64]),
    "f2_strings_export": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, i64p, C.c_int64]),
    "f2_strings_build_lists": (C.c_int, [C.c_void_p, C.c_int]),
    "f2_strings_determinants": (C.c_int, [C.c_void_p, u64p]),
    "f2_strings_determinant": (C.c_int, [C.c_void_p, C.c_int64, u64p]),
    "f2_strings_determinant_index": (C.c_int64, [C.c_void_p, C.c_uint64, C.c_uint64]),
    "f2_sigma_create": (C.c_void_p, [C.c_void_p, C.c_double, f64p, f64p, C.c_int, C.c_int]),
    "f2_sigma_destroy": (None, [C.c_void_p]),
    "f2_sigma_set_hamiltonian": (C.c_int, [C.c_void_p, C.c_double, f64p, f64p]),
    "f2_sigma_set_algorithm": (C.c_int, [C.c_void_p, C.c_char_p]),
    "f2_sigma_get_algorithm": (C.c_char_p, [C.c_void_p]),
    "f2_sigma_set_memory": (C.c_int, [C.c_void_p, C.c_int64]),
    "f2_sigma_set_log_level": (C.c_int, [C.c_void_p, C.c_int]),
    "f2_sigma_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "f2_sigma_apply_host": (C.c_int, [C.c_void_p, f64p, f64p, C.c_int64]),
    "f2_sigma_times": (C.c_int, [C.c_void_p, f64p]),
    "f2_sigma_last_ms": (C.c_int, [C.c_void_p, f64p]),
    "f2_sigma_hdiag": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_sigma_rdm": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_sigma_rdm_part": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "f2_sigma_rdm_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_sigma_trdm1": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_dist_unique_id": (C.c_int, [C.c_char_p]),
    "f2_dist_init": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_int]),
    "f2_dist_local_size": (C.c_int, [C.c_void_p, i64p]),
    "f2_dist_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_dist_allreduce": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "f2_dist_finalize": (C.c_int, [C.c_void_p]),
    "f2_sigma_launch_count": (C.c_int64, [C.c_void_p]),
    "f2_sigma_t3_shape": (C.c_int, [C.c_void_p, i64p]),
    "f2_sigma_part_layout": (C.c_int, [C.c_void_p, C.c_int, C.c_int, i64p, i64p]),
    "f2_sigma_apply_part": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                      C.c_void_p]),
    "f2_sigma_hdiag_part": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "f2_sigma_hdiag_spin_averaged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "f2_sigma_apply_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p]),
    "f2_spin2_matrix": (C.c_int, [C.c_void_p, i64p, C.c_int, f64p]),
    "f2_sigma_rows_form": (C.c_int, [C.c_void_p, C.c_int, C.c_int, i64p]),
    "f2_sigma_apply_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "f2_sigma_apply_rows_t3_piece": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                               C.c_int, C.c_void_p]),
    "f2_slater_matrix": (C.c_int, [C.c_void_p, C.c_double, f64p, f64p, i64p, C.c_int, f64p]),
    "f2_adapter_create": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "f2_adapter_destroy": (None, [C.c_void_p]),
    "f2_adapter_info": (C.c_int, [C.c_void_p, i64p]),
    "f2_adapter_table": (C.c_int64, [C.c_void_p, i64p, i64p, f64p, C.c_int64]),
    "f2_adapter_conf_ranges": (C.c_int, [C.c_void_p, i64p]),
    "f2_adapter_csf_to_det": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "f2_adapter_det_to_csf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "f2_adapter_csf_to_det_host": (C.c_int, [C.c_void_p, f64p, C.c_int64, f64p, C.c_int64]),
    "f2_adapter_det_to_csf_host": (C.c_int, [C.c_void_p, f64p, C.c_int64, f64p, C.c_int64]),
    "f2_adapter_hdiag_csf": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "f2_dl_workspace_doubles": (C.c_int64, []),
    "f2_dl_dots": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, C.c_int64,
                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_dl_combine": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, f64p, C.c_void_p, C.c_int64, C.c_int,
                                C.c_double, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "f2_dl_residual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, f64p, f64p, C.c_void_p,
                                 C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int64, C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "f2_dl_scale": (C.c_int, [C.c_void_p, C.c_double, C.c_int64, C.c_void_p]),
    "f2_dl_residual_fused": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, f64p, f64p, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "f2_dl_project": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int, f64p, f64p,
                                C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import "
                "__graft_entry__ as g; g.build()' or make -C forte2_b200/csrc). forte2_b200 has no "
                "CPU fallback.")
        # F2_LIB_PATH: another build of the same library (kernel tuning variants, scripts/tune_ssg.sh)
        L = C.CDLL(os.environ.get("F2_LIB_PATH") or str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return lib().f2_last_error().decode()


def check(status: int):
    if status:
        raise _EXC.get(status, RuntimeError)(last_error())
