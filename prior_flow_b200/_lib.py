"""ctypes binding of libpriorcorr.so (include/priorcorr.h).

The library is loaded lazily and the product fails loudly when it is missing: there is no CPU or
PyTorch fallback for any entry point.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libpriorcorr.so")
ABI_VERSION = 5
MAX_LEVELS = 4

DIV_IEEE, DIV_ATEN_CUDA = 0, 1
VOL_FP32_3XF16, VOL_F16, VOL_FP32_SIMT = 0, 1, 2
VOLUME_MODES = {"fp32": VOL_FP32_3XF16, "fp32_3xf16": VOL_FP32_3XF16, "f16": VOL_F16, "fp32_simt": VOL_FP32_SIMT}

_fp = C.c_void_p  # device pointers travel as integers
_LevelPtrs = _fp * MAX_LEVELS


class VolumeArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("channels", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("num_levels", C.c_int), ("mode", C.c_int),
                ("fmap1", _fp), ("fmap2", _fp), ("level", _LevelPtrs),
                ("workspace", _fp), ("workspace_bytes", C.c_longlong)]


class LookupArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("h", C.c_int), ("w", C.c_int), ("h2", C.c_int), ("w2", C.c_int),
                ("radius", C.c_int), ("num_levels", C.c_int), ("cyclic", C.c_int), ("div_mode", C.c_int),
                ("coords", _fp), ("own", _LevelPtrs), ("other", _LevelPtrs),
                ("grid_w2c", _fp), ("grid_c2w", _fp), ("grid_batch_stride", C.c_longlong),
                ("out_own", _fp), ("out_other", _fp), ("scratch", _fp),
                ("dbg_own_xy", _fp), ("dbg_other_xy", _fp),
                ("out_channels_last", C.c_int), ("fuse_sum", C.c_int), ("scratch_own", _fp), ("no_rotate", C.c_int)]


class DcclConvArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("h", C.c_int), ("w", C.c_int), ("in_channels", C.c_int), ("out_channels", C.c_int),
                ("div_mode", C.c_int), ("split", C.c_int), ("out_channels_last", C.c_int), ("after_lookup", C.c_int),
                ("raw", _fp), ("own_cl", _fp), ("grid_c2w", _fp), ("grid_batch_stride", C.c_longlong),
                ("prepared_weight", _fp), ("bias", _fp), ("out", _fp)]


class VolumeBwdArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("channels", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("fmap1", _fp), ("fmap2", _fp), ("dvolume", _fp), ("dfmap1", _fp), ("dfmap2", _fp),
                ("workspace", _fp), ("workspace_bytes", C.c_longlong),
                ("query_begin", C.c_int), ("query_count", C.c_int), ("accumulate_dfmap2", C.c_int), ("planes_ready", C.c_int)]


class OnTheFlyArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("channels", C.c_int), ("h", C.c_int), ("w", C.c_int),
                ("radius", C.c_int), ("num_levels", C.c_int), ("cyclic", C.c_int), ("div_mode", C.c_int),
                ("coords", _fp), ("fmap1_own", _fp), ("fmap2_own", _LevelPtrs),
                ("fmap1_other", _fp), ("fmap2_other", _LevelPtrs),
                ("grid_w2c", _fp), ("grid_c2w", _fp), ("grid_batch_stride", C.c_longlong),
                ("out_own", _fp), ("out_other", _fp), ("scratch", _fp)]


class OnTheFlyTcArgs(C.Structure):
    _fields_ = [("base", OnTheFlyArgs),
                ("f1_hi_own", _fp), ("f1_lo_own", _fp), ("f2_hi_own", _LevelPtrs), ("f2_lo_own", _LevelPtrs),
                ("f1_hi_other", _fp), ("f1_lo_other", _fp), ("f2_hi_other", _LevelPtrs), ("f2_lo_other", _LevelPtrs),
                ("amax_own", _fp), ("amax_other", _fp), ("pool", _fp), ("pool_segments", C.c_longlong), ("worklist", _fp), ("tap_xy", _fp), ("no_rotate", C.c_int)]


class RemapArgs(C.Structure):
    _fields_ = [("batch", C.c_int), ("channels", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("Ho", C.c_int), ("Wo", C.c_int), ("cyclic", C.c_int), ("div_mode", C.c_int),
                ("src", _fp), ("coords", _fp),
                ("coord_batch_stride", C.c_longlong), ("coord_pixel_stride", C.c_longlong),
                ("coord_xy_stride", C.c_longlong), ("out", _fp)]


class LookupBwdArgs(C.Structure):
    _fields_ = [("fwd", LookupArgs), ("grad_own", _fp), ("grad_other", _fp),
                ("dgrad_own", _LevelPtrs), ("dgrad_other", _LevelPtrs),
                ("query_begin", C.c_int), ("query_count", C.c_int), ("scratch_ready", C.c_int)]


# name -> (restype, argtypes); the single source of truth for tests/test_abi.py as well
SIGNATURES = {
    "pf_abi_version": (C.c_int, []),
    "pf_last_error": (C.c_char_p, []),
    "pf_build_info": (C.c_char_p, []),
    "pf_volume_workspace_bytes": (C.c_longlong, [C.c_int] * 5),
    "pf_volume_build": (C.c_int, [C.POINTER(VolumeArgs), _fp]),
    "pf_avg_pool2x2": (C.c_int, [_fp, _fp, C.c_longlong, C.c_int, C.c_int, _fp]),
    "pf_lookup_dual": (C.c_int, [C.POINTER(LookupArgs), _fp]),
    "pf_lookup_onthefly": (C.c_int, [C.POINTER(OnTheFlyArgs), _fp]),
    "pf_lookup_onthefly_tc": (C.c_int, [C.POINTER(OnTheFlyTcArgs), _fp]),
    "pf_onthefly_absmax": (C.c_int, [_fp, C.c_longlong, _fp, _fp]),
    "pf_onthefly_split": (C.c_int, [_fp, C.c_longlong, _fp, _fp, _fp, C.c_longlong, _fp]),
    "pf_fmap_pyramid": (C.c_int, [_fp, C.POINTER(_fp), C.c_int, C.c_longlong, C.c_int, C.c_int, _fp]),
    "pf_samplegrid": (C.c_int, [_fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, _fp]),
    "pf_remap": (C.c_int, [C.POINTER(RemapArgs), _fp]),
    "pf_flo_rotate": (C.c_int, [_fp, _fp, _fp, C.c_longlong, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "pf_warp_groupcorr": (C.c_int, [_fp, _fp, _fp, _fp] + [C.c_int] * 6 + [_fp]),
    "pf_lookup_dual_bwd": (C.c_int, [C.POINTER(LookupBwdArgs), _fp]),
    "pf_remap_bwd": (C.c_int, [C.POINTER(RemapArgs), _fp, _fp, _fp]),
    "pf_pyramid_fold_bwd": (C.c_int, [C.POINTER(_fp), C.c_int, C.c_longlong, C.c_int, C.c_int, _fp]),
    "pf_warp_groupcorr_bwd": (C.c_int, [_fp] * 6 + [C.c_int] * 6 + [_fp]),
    "pf_dccl_conv_weight_bytes": (C.c_longlong, []),
    "pf_dccl_conv_prepare": (C.c_int, [_fp, C.c_int, C.c_int, _fp, _fp]),
    "pf_dccl_conv": (C.c_int, [C.POINTER(DcclConvArgs), _fp]),
    "pf_volume_bwd_workspace_bytes": (C.c_longlong, [C.c_int] * 4),
    "pf_volume_bwd": (C.c_int, [C.POINTER(VolumeBwdArgs), _fp]),
    "pf_convex_upsample": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "pf_convex_upsample_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_int, _fp]),
    "pf_uniform_loss_fwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_float, C.c_int, C.c_int, C.c_int, _fp]),
    "pf_uniform_loss_bwd": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_float, _fp, C.c_int, C.c_int, C.c_int, _fp]),
    "pf_great_circle": (C.c_int, [_fp, _fp, _fp, C.c_int, C.c_int, C.c_int, C.c_float, _fp]),
    "pf_probe_gather": (C.c_int, [_fp, C.c_longlong, C.c_int, C.c_int, _fp, _fp, _fp]),
    "pf_probe_stream_read": (C.c_int, [_fp, C.c_longlong, _fp, _fp]),
}

_lib = None
_lock = threading.Lock()


class PriorCorrError(RuntimeError):
    pass


def load():
    """Loads libpriorcorr.so (once).  Raises if it is absent or has the wrong ABI — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise PriorCorrError(
                f"{LIB_PATH} is missing: build it with `python -m prior_flow_b200.build` "
                "(nvcc, sm_100a). prior_flow_b200 has no CPU / PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        if lib.pf_abi_version() != ABI_VERSION:
            raise PriorCorrError(f"libpriorcorr ABI {lib.pf_abi_version()} != expected {ABI_VERSION}")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        raise PriorCorrError(f"{what}: {load().pf_last_error().decode()}")


def level_ptrs(tensors):
    arr = _LevelPtrs()
    for i in range(MAX_LEVELS):
        arr[i] = tensors[i].data_ptr() if i < len(tensors) and tensors[i] is not None else None
    return arr
