"""ctypes binding of the C-ABI declared in include/stylemesh_b200.h.

There is deliberately NO fallback: if the shared library is missing or a call fails, this module raises.
(The CPU oracle under oracle/ is test infrastructure and is never imported from here.)
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG_DIR, "lib", "libstylemesh_b200.so")
if os.environ.get("SMB_LIB"):            # kernel experiments: an alternative build of the same ABI (tools/build_variant.sh)
    LIB_PATH = os.environ["SMB_LIB"]

NUM_VGG_CONVS = 13
IMPL_SIMT = 0
IMPL_TC = 1
IMPL_TC_PH = 5

# (name, restype, argtypes) — must list every symbol of include/stylemesh_b200.h (checked by tests/test_abi.py)
_f = C.c_float
_i = C.c_int
_p = C.c_void_p
_i64 = C.c_int64
_d = C.c_double
_pp = C.POINTER(C.c_void_p)
_ip = C.POINTER(C.c_int)

PROTOTYPES = [
    ("smb_abi_version", _i, []),
    ("smb_last_error", C.c_char_p, []),
    ("smb_uv_sample_fwd", _i, [_pp, _ip, _ip, _i, _i, _p, _i, _i, _f, _f, _p, _p]),
    ("smb_uv_texel_index", _i, [_p, _i, _i, _i, _p, _p, _p]),
    ("smb_uv_scatter_bwd", _i, [_pp, _ip, _ip, _i, _i, _p, _i, _i, _p, _p, _p, _p]),
    ("smb_adam_step", _i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _i, _f, _f, _f, _f, _p]),
    ("smb_texreg_value", _i, [_p, _i64, _f, _f, _f, _p, _p]),
    ("smb_adam_step_segments", _i, [_p, _p, _p, _p, _i64, _p, _p, _i, _f, _f, _f, _f, _i, _f, _f, _f, _p]),
    ("smb_texreg_value_segments", _i, [_p, _i64, _p, _p, _i, _f, _f, _p, _p]),
    ("smb_dist_adam_step", _i, [_i, _i, _p, _p, _p, _p, _p, _i64, _p, _p, _i, _f, _f, _f, _f, _i, _f, _f, C.c_uint, _p]),
    ("smb_view_uv_to_grid", _i, [_p, _i, _i, _p, _p, _p, _p]),
    ("smb_view_gather2d", _i, [_p, _i, _i, _i, _p, _p, _i, _i, _p, _p]),
    ("smb_view_resize_linear", _i, [_p, _i, _d, _i, _i, _p, _p, _p, _p, _i, _i, _p, _p]),
    ("smb_view_depth_levels", _i, [_p, _i64, _p, _i, _d, _i, _p, _p, _p, _p, _p, _p]),
    ("smb_view_rgb_pre", _i, [_p, _i, _i, _p, _p]),
    ("smb_view_angle_degrees", _i, [_p, _i64, _p, _p]),
    ("smb_view_erode3x3", _i, [_p, _i, _i, _p, _p]),
    ("smb_view_level_masks", _i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p]),
    ("smb_view_level_plan", _i, [_p, _p, _p, _p, _f, _i, _i, _i, _i, _p, _p, _i, _ip, _ip, _p, _i, _p, _p]),
    ("smb_texture_post_rgb8", _i, [_p, _i, _i, _p, _p]),
    ("smb_mip_downsample2x", _i, [_p, _i, _i, _p, _p]),
    ("smb_mip_preview", _i, [_pp, _ip, _ip, _i, _p, _i, _i, _i, _f, _p, _p]),
    ("smb_raster_view", _i, [_p, _i, _p, _i, _p, _p, _p, _p, _i, _i, _f, _f, _f, _i, _p, _p, _p, _p, _p, _p]),
    ("smb_ctx_create", _p, []),
    ("smb_ctx_destroy", None, [_p]),
    ("smb_ctx_set_impl", _i, [_p, _i, _i]),
    ("smb_ctx_load_vgg", _i, [_p, _pp, _pp, _i]),
    ("smb_level_begin", _i, [_p, _i, _i]),
    ("smb_level_forward", _i, [_p, _i, _p, _i, _p]),
    ("smb_level_forward_features", _i, [_p, _i, _p, _i, C.c_uint, _p]),
    ("smb_level_feature_shape", _i, [_p, _i, _i, _ip, _ip, _ip]),
    ("smb_level_get_feature", _i, [_p, _i, _i, _p, _p]),
    ("smb_level_get_feature_nhwc", _i, [_p, _i, _i, _p, _p]),
    ("smb_ctx_release_slots", _i, [_p]),
    ("smb_level_gram", _i, [_p, _i, _i, _p, _f, _p, _p]),
    ("smb_level_style_term", _i, [_p, _i, _i, _p, _f, _p, _f, _p, _f, _p, _f, _p, _p, _p]),
    ("smb_level_content_term", _i, [_p, _i, _i, _p, _p, _f, _f, _p, _p]),
    ("smb_level_backward", _i, [_p, _i, _p, _p]),
    ("smb_launch_count", _i64, []),
    ("smb_ctx_set_timing", _i, [_p, _i]),
    ("smb_ctx_read_timing", _i, [_p, _p, _p, _p, _i]),
    ("smb_ctx_device_bytes", _i64, [_p]),
    ("smb_debug_set_igemm_trace", _i, [_p]),
    ("smb_unit_conv3x3", _i, [_i, _p, _i, _i, _i, _p, _p, _i, _i, _i, _p, _p]),
    ("smb_unit_conv3x3_fused", _i, [_p, _i, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    ("smb_unit_maxpool", _i, [_p, _i, _i, _i, _p, _p]),
    ("smb_unit_maxpool_bwd", _i, [_p, _p, _i, _i, _i, _p, _p]),
    ("smb_unit_gram", _i, [_i, _p, _i, _i, _i, _p, _f, _p, _p]),
]

_lib = None
_lock = threading.Lock()


class StyleMeshB200Error(RuntimeError):
    """A call into libstylemesh_b200.so failed."""


def load():
    """Load the shared library (once).  Raises if it has not been built — there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise StyleMeshB200Error(
                f"{LIB_PATH} not found: build it with `python -m stylemesh_b200.build` "
                "(or __graft_entry__.build()). stylemesh_b200 has no CPU/PyTorch fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in PROTOTYPES:
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.smb_abi_version() != 1:
            raise StyleMeshB200Error(f"ABI version mismatch: library reports {lib.smb_abi_version()}, binding expects 1")
        _lib = lib
        return _lib


def last_error() -> str:
    msg = load().smb_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str) -> int:
    if rc < 0:
        raise StyleMeshB200Error(f"{what} failed (status {rc}): {last_error()}")
    return rc


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (C.c_void_p * len(tensors))()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr()
    return arr


def int_array(values):
    return (C.c_int * len(values))(*[int(v) for v in values])


_raw_stream = None


def current_stream():
    """torch's current CUDA stream of the current device as a cudaStream_t.  Called once per C-ABI call (~70 per
    step): goes through torch._C's raw accessors (sub-microsecond) instead of building torch.cuda.Stream objects
    (~13 us each, a quarter of the host time of a step)."""
    global _raw_stream
    import torch
    if _raw_stream is None:
        get_raw, get_dev = getattr(torch._C, "_cuda_getCurrentRawStream", None), getattr(torch._C, "_cuda_getDevice", None)
        if get_raw is not None and get_dev is not None:
            _raw_stream = lambda: get_raw(get_dev())
        else:                                                   # pragma: no cover - older torch
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream
    return C.c_void_p(_raw_stream())
