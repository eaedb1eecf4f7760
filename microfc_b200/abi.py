"""ctypes binding of include/mfc_b200.h -- the same stubs a reference maintainer would write
in Fortran with ISO_C_BINDING (fortran/m_b200_bindings.f90).

The library is loaded from the package directory (built in-tree by __graft_entry__.build()).
If it is missing, or no CUDA device is usable, every call FAILS LOUDLY: there is no CPU
fallback (the CPU oracle under oracle/ is test infrastructure and is never used here).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

MAX_FLUIDS = 4
ABI_VERSION = 1
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MFC_B200_LIB") or os.path.join(_HERE, "libmfc_b200.so")

c_double_p = C.POINTER(C.c_double)


class Params(C.Structure):
    """mfc_b200_params_t"""
    _fields_ = [
        ("abi_version", C.c_int32),
        ("m", C.c_int32), ("n", C.c_int32), ("p", C.c_int32),
        ("m_glb", C.c_int32), ("n_glb", C.c_int32), ("p_glb", C.c_int32),
        ("num_dims", C.c_int32), ("num_fluids", C.c_int32), ("sys_size", C.c_int32), ("buff_size", C.c_int32),
        ("weno_order", C.c_int32), ("weno_eps", C.c_double),
        ("time_stepper", C.c_int32), ("weno_Re_flux", C.c_int32), ("run_time_info", C.c_int32),
        ("t_step_start", C.c_int32), ("t_step_stop", C.c_int32),
        ("bc", C.c_int32 * 6),
        ("proc_rank", C.c_int32), ("num_procs", C.c_int32),
        ("proc_coords", C.c_int32 * 3), ("num_procs_dir", C.c_int32 * 3),
        ("gammas", C.c_double * MAX_FLUIDS), ("pi_infs", C.c_double * MAX_FLUIDS),
        ("Re", (C.c_double * 2) * MAX_FLUIDS),
        ("cb", c_double_p * 3), ("cc", c_double_p * 3), ("ds", c_double_p * 3),
        ("strict_math", C.c_int32), ("device", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


MAX_PATCHES = 10


class Patch(C.Structure):
    """mfc_b200_patch_t (patch_icpp(i), src/pre_process/m_global_parameters.fpp:196-230)"""
    _fields_ = [
        ("geometry", C.c_int32), ("smoothen", C.c_int32), ("smooth_patch_id", C.c_int32),
        ("alter_patch", C.c_int32 * (MAX_PATCHES + 1)),
        ("x_centroid", C.c_double), ("y_centroid", C.c_double), ("z_centroid", C.c_double),
        ("length_x", C.c_double), ("length_y", C.c_double), ("length_z", C.c_double),
        ("radius", C.c_double), ("radii", C.c_double * 3), ("normal", C.c_double * 3),
        ("epsilon", C.c_double), ("smooth_coeff", C.c_double),
        ("vel", C.c_double * 3), ("pres", C.c_double),
        ("alpha_rho", C.c_double * MAX_FLUIDS), ("alpha", C.c_double * MAX_FLUIDS),
    ]


class MfcB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmfc_b200: {msg} (code {code})")
        self.code = code


def field_pointers(arrs: Sequence[np.ndarray]):
    """Array of base pointers, one per field (what c_loc(q(i)%sf) gives the Fortran host)."""
    ptrs = (c_double_p * len(arrs))()
    for i, a in enumerate(arrs):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        ptrs[i] = a.ctypes.data_as(c_double_p)
    return ptrs


_lib = None


def lib() -> C.CDLL:
    """Load libmfc_b200.so; raise if it is not built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MfcB200Error(-2, f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(the hot path is CUDA-only, there is no CPU fallback)")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    pp = C.POINTER(c_double_p)
    L.mfc_b200_init.argtypes = [C.POINTER(Params)]
    L.mfc_b200_get_unique_id.argtypes = [C.c_char_p]
    L.mfc_b200_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int]
    L.mfc_b200_upload.argtypes = [pp]
    L.mfc_b200_step.argtypes = [C.c_int, C.c_double, c_double_p, c_double_p]
    L.mfc_b200_step_async.argtypes = [C.c_int, C.c_double, C.c_int]
    L.mfc_b200_sync.argtypes = []
    L.mfc_b200_compute_rhs.argtypes = [pp, pp]
    L.mfc_b200_download.argtypes = [pp]
    L.mfc_b200_generate_initial_condition.argtypes = [C.c_int32, C.POINTER(Patch), pp, C.c_double]
    L.mfc_b200_generate_initial_condition2.argtypes = [C.c_int32, C.POINTER(Patch), pp, pp, C.c_double]
    L.mfc_b200_download_prim.argtypes = [pp]
    L.mfc_b200_finalize.argtypes = []
    L.mfc_b200_last_error.restype = C.c_char_p
    L.mfc_b200_get_weno_coefficients.argtypes = [C.c_int] + [c_double_p] * 5
    L.mfc_b200_kernel_launches.restype = C.c_int64
    L.mfc_b200_state_snapshot.argtypes = []
    L.mfc_b200_state_restore.argtypes = []
    L.mfc_b200_timer_start.argtypes = []
    L.mfc_b200_timer_stop.argtypes = [c_double_p]
    L.mfc_b200_profile_enable.argtypes = [C.c_int]
    L.mfc_b200_profile_get.argtypes = [C.c_int, c_double_p, C.POINTER(C.c_int64)]
    L.mfc_b200_kernel_name.argtypes = [C.c_int]
    L.mfc_b200_kernel_name.restype = C.c_char_p
    L.mfc_b200_debug_fill_ghosts.argtypes = [C.c_int]
    _lib = L
    return L


def check(code: int) -> None:
    if code != 0:
        msg = lib().mfc_b200_last_error()
        raise MfcB200Error(code, msg.decode() if msg else "unknown error")


EXPORTED_SYMBOLS = [
    "mfc_b200_init", "mfc_b200_get_unique_id", "mfc_b200_comm_init", "mfc_b200_upload",
    "mfc_b200_step", "mfc_b200_step_async", "mfc_b200_sync", "mfc_b200_compute_rhs",
    "mfc_b200_download", "mfc_b200_download_prim", "mfc_b200_finalize", "mfc_b200_last_error",
    "mfc_b200_get_weno_coefficients", "mfc_b200_kernel_launches", "mfc_b200_state_snapshot",
    "mfc_b200_state_restore", "mfc_b200_timer_start", "mfc_b200_timer_stop", "mfc_b200_profile_enable", "mfc_b200_profile_get", "mfc_b200_kernel_name",
    "mfc_b200_generate_initial_condition", "mfc_b200_generate_initial_condition2", "mfc_b200_debug_fill_ghosts",
]
