"""ctypes binding of libdfk.so (the C ABI declared in include/dfk.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C deepfactors_b200/csrc`.
There is NO fallback: if the shared library is missing or a call fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFK_LIB: another build of the same library (A/B measurements of kernel variants); default = the in-tree build
LIB_PATH = os.environ.get("DFK_LIB") or os.path.join(_HERE, "libdfk.so")

DFK_OK = 0
DFK_ERR_INVALID_ARG = 1
DFK_ERR_CUDA = 2
DFK_ERR_UNSUPPORTED = 3
DFK_ERR_NOMEM = 4

DFK_GRAM_AUTO, DFK_GRAM_FP32, DFK_GRAM_TF32X3 = 0, 1, 2


class DfkError(RuntimeError):
    """std::runtime_error / vc::CUDAException stand-in carrying the status code."""

    def __init__(self, status: int, message: str):
        super().__init__(f"dfk status {status}: {message}")
        self.status = status
        self.message = message


class DfkImage(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("pitch_bytes", C.c_size_t), ("width", C.c_uint32), ("height", C.c_uint32)]


class DfkCamera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("u0", C.c_float), ("v0", C.c_float),
                ("width", C.c_float), ("height", C.c_float)]


class DfkDenseSfmParams(C.Structure):
    _fields_ = [("huber_delta", C.c_float), ("ocl_th", C.c_float), ("avg_dpt", C.c_float),
                ("min_dpt", C.c_float), ("valid_border", C.c_int32)]


class DfkSfmAlignerParams(C.Structure):
    _fields_ = [("sfmparams", DfkDenseSfmParams), ("step_threads", C.c_int32), ("step_blocks", C.c_int32),
                ("eval_threads", C.c_int32), ("eval_blocks", C.c_int32)]


class DfkSfmWorkItem(C.Structure):
    _fields_ = [("pose0", C.c_float * 7), ("pose1", C.c_float * 7), ("cam", DfkCamera),
                ("img0", DfkImage), ("img1", DfkImage), ("dpt0", DfkImage), ("valid0", DfkImage),
                ("prx0_jac", DfkImage), ("grad1", DfkImage),
                ("prx_orig", DfkImage), ("code", C.POINTER(C.c_float))]  # optional fused depth decode


class DfkTrackLevel(C.Structure):
    _fields_ = [("cam", DfkCamera), ("img0", DfkImage), ("img1", DfkImage), ("dpt0", DfkImage), ("grad1", DfkImage),
                ("iterations", C.c_int)]


class DfkWindowDesc(C.Structure):
    _fields_ = [("num_keyframes", C.c_int32), ("num_pairs", C.c_int32), ("num_items", C.c_int32), ("code_size", C.c_int32),
                ("pair_k0", C.POINTER(C.c_int32)), ("pair_k1", C.POINTER(C.c_int32)), ("item_pair", C.POINTER(C.c_int32)),
                ("item_width", C.POINTER(C.c_int32)), ("item_height", C.POINTER(C.c_int32))]


# every symbol include/dfk.h declares: (name, restype, argtypes)
_F = C.POINTER(C.c_float)
_IMG = C.POINTER(DfkImage)
_CAM = C.POINTER(DfkCamera)
_H = C.c_void_p
SYMBOLS = {
    "dfk_create": (C.c_int, [C.c_int, C.POINTER(_H)]),
    "dfk_destroy": (C.c_int, [_H]),
    "dfk_set_stream": (C.c_int, [_H, C.c_void_p]),
    "dfk_set_sm_limit": (C.c_int, [_H, C.c_int]),
    "dfk_use_own_stream": (C.c_int, [_H]),
    "dfk_get_stream": (C.c_void_p, [_H]),
    "dfk_synchronize": (C.c_int, [_H]),
    "dfk_last_error": (C.c_char_p, [_H]),
    "dfk_status_string": (C.c_char_p, [C.c_int]),
    "dfk_version": (C.c_int, []),
    "dfk_sfm_supports_code_size": (C.c_int, [C.c_int]),
    "dfk_set_profiling": (C.c_int, [_H, C.c_int]),
    "dfk_get_profile": (C.c_int, [_H, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "dfk_sfm_set_params": (C.c_int, [_H, C.POINTER(DfkSfmAlignerParams)]),
    "dfk_sfm_get_params": (C.c_int, [_H, C.POINTER(DfkSfmAlignerParams)]),
    "dfk_sfm_set_gram_mode": (C.c_int, [_H, C.c_int]),
    "dfk_se3_set_huber_delta": (C.c_int, [_H, C.c_float]),
    "dfk_sfm_run_step": (C.c_int, [_H, _F, _F, _F, C.c_int, _CAM, _IMG, _IMG, _IMG, _IMG, _IMG, _IMG, _IMG,
                                   _F, _F, _F, C.POINTER(C.c_uint64)]),
    "dfk_sfm_evaluate_error": (C.c_int, [_H, _F, _F, _CAM, _IMG, _IMG, _IMG, _IMG, _IMG, _F,
                                         C.POINTER(C.c_uint64)]),
    "dfk_sfm_run_step_batch": (C.c_int, [_H, C.POINTER(DfkSfmWorkItem), C.c_int, C.c_int, C.c_void_p]),
    "dfk_sfm_run_step_batch_host": (C.c_int, [_H, C.POINTER(DfkSfmWorkItem), C.c_int, C.c_int, _F]),
    "dfk_sfm_stream_create": (C.c_int, [_H, C.c_int, C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "dfk_sfm_stream_destroy": (C.c_int, [_H, C.c_void_p]),
    "dfk_sfm_stream_submit": (C.c_int, [_H, C.c_void_p, C.POINTER(DfkSfmWorkItem), C.c_int, C.POINTER(C.c_uint64)]),
    "dfk_sfm_stream_wait": (C.c_int, [_H, C.c_void_p, C.c_uint64, _F]),
    "dfk_window_create": (C.c_int, [_H, C.POINTER(DfkWindowDesc), C.POINTER(C.c_void_p)]),
    "dfk_window_destroy": (C.c_int, [_H, C.c_void_p]),
    "dfk_window_floats": (C.c_size_t, [C.c_void_p]),
    "dfk_window_assemble": (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dfk_se3_run_step": (C.c_int, [_H, _F, _CAM, _IMG, _IMG, _IMG, _IMG, _F, _F, _F, C.POINTER(C.c_uint64)]),
    "dfk_se3_track": (C.c_int, [_H, _F, C.POINTER(DfkTrackLevel), C.c_int, _F, _F, _F, _F, C.c_int]),
    "dfk_se3_warp": (C.c_int, [_H, _F, _CAM, _IMG, _IMG, _IMG, _IMG, _F, C.POINTER(C.c_uint64)]),
    "dfk_depth_run_step": (C.c_int, [_H, _F, C.c_int, _IMG, _IMG, _IMG, _F, _F, _F, C.POINTER(C.c_uint64)]),
    "dfk_reprojection_linearize": (C.c_int, [_H, _F, _F, _F, C.c_int, _CAM, _IMG, _IMG, C.c_int, _F, _F, C.c_float,
                                             C.c_float, _F, _F]),
    "dfk_sparse_geometric_linearize": (C.c_int, [_H, _F, _F, _F, _F, C.c_int, _CAM, _IMG, _IMG, _IMG, _IMG, _IMG, C.c_int,
                                                 C.POINTER(C.c_int), C.c_float, _F, C.POINTER(C.c_int)]),
    "dfk_update_depth": (C.c_int, [_H, _F, C.c_int, _IMG, _IMG, C.c_float, _IMG]),
    "dfk_sobel_gradients": (C.c_int, [_H, _IMG, _IMG]),
    "dfk_gaussian_blur_down": (C.c_int, [_H, _IMG, _IMG]),
    "dfk_build_image_pyramid": (C.c_int, [_H, _IMG, _IMG, C.c_int]),
    "dfk_squared_error": (C.c_int, [_H, _IMG, _IMG, _F]),
}

_lib = None


def lib():
    """Load libdfk.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; "
                "g.build()' or make -C deepfactors_b200/csrc). deepfactors_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(handle, status: int):
    if status != DFK_OK:
        msg = lib().dfk_last_error(handle)
        raise DfkError(status, (msg or b"").decode() or lib().dfk_status_string(status).decode())


def record_floats(code_size: int) -> int:
    npar = 12 + code_size
    return npar * (npar + 1) // 2 + npar + 2
