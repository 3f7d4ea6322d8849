"""ctypes binding of the C ABI in include/bsvd_b200.h.

This is the stub a maintainer of the reference would add next to
Experimental_root/archs/bsvd_arch.py to call the B200 path.  It only
marshals pointers and sizes; torch is used by callers for device memory and streams, never for
the arithmetic.  If the shared library is missing the import fails loudly — there is no
fallback implementation.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSVD_B200_LIB: development override (A/B timing of two builds on the same box)
LIB_PATH = os.environ.get("BSVD_B200_LIB") or os.path.join(_HERE, "lib", "libbsvd_b200.so")

NUM_LAYERS = 32
PREC_FP16, PREC_BF16, PREC_FP32X3 = 0, 1, 2
EPI_RELU6, EPI_PIXSHUF, EPI_SKIP_ADD, EPI_SHIFT_STORE, EPI_STRIDE2 = 1, 2, 4, 8, 16

# every symbol include/bsvd_b200.h declares (tests check that the library exports all of them)
EXPORTS = (
    "bsvd_create", "bsvd_destroy", "bsvd_last_error", "bsvd_version", "bsvd_set_weights",
    "bsvd_layer_shape", "bsvd_forward_clip", "bsvd_forward_clips", "bsvd_forward_clip_host", "bsvd_stream_push",
    "bsvd_reset", "bsvd_last_launch_count", "bsvd_workspace_bytes", "bsvd_conv_stage",
    "bsvd_set_profiling", "bsvd_get_stage_ms", "bsvd_stage_info", "bsvd_last_stage_ms",
    "bsvd_forward_clip_host_async", "bsvd_host_sync", "bsvd_denoise_clip", "bsvd_psnr",
    "bsvd_denoise_clip_u8", "bsvd_overflow_flag", "bsvd_host_last_output", "bsvd_stream_graph_replays", "bsvd_ssim",
    "bsvd_peer_create", "bsvd_peer_handle_bytes", "bsvd_peer_get_handle", "bsvd_peer_open",
    "bsvd_peer_local_data", "bsvd_peer_put", "bsvd_peer_put2d", "bsvd_peer_put3d", "bsvd_peer_signal", "bsvd_peer_wait",
    "bsvd_peer_read_flag", "bsvd_peer_destroy",
)
NUM_STAGES = 33


class BsvdConfig(C.Structure):
    _fields_ = [("chns", C.c_int * 3), ("mid_ch", C.c_int), ("interm_ch", C.c_int),
                ("in_ch", C.c_int), ("out_ch", C.c_int), ("act_relu6", C.c_int),
                ("norm_none", C.c_int), ("precision", C.c_int), ("device", C.c_int)]


class BsvdConvDesc(C.Structure):
    _fields_ = [("T", C.c_int), ("H", C.c_int), ("W", C.c_int), ("cin", C.c_int),
                ("cout", C.c_int), ("flags", C.c_int), ("precision", C.c_int),
                ("debug_variant", C.c_int)]


class BsvdError(RuntimeError):
    pass


_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libbsvd_b200.so (built by __graft_entry__.build()).  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise BsvdError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The B200 path has no CPU fallback.")
    lib = C.CDLL(p)
    vp, ci, cip = C.c_void_p, C.c_int, C.POINTER(C.c_int)
    lib.bsvd_last_error.restype = C.c_char_p
    lib.bsvd_version.restype = C.c_char_p
    lib.bsvd_create.argtypes = [C.POINTER(BsvdConfig), C.POINTER(vp)]
    lib.bsvd_destroy.argtypes = [vp]
    lib.bsvd_set_weights.argtypes = [vp, ci, vp, vp, ci, ci]
    lib.bsvd_layer_shape.argtypes = [vp, ci, cip, cip, cip]
    lib.bsvd_forward_clip.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.bsvd_forward_clips.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, vp]
    lib.bsvd_forward_clip_host.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.bsvd_forward_clip_host_async.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, vp]
    lib.bsvd_host_sync.argtypes = [vp]
    lib.bsvd_host_last_output.argtypes = [vp, C.POINTER(vp)]
    lib.bsvd_denoise_clip.argtypes = [vp, vp, C.c_float, vp, ci, ci, ci, vp]
    lib.bsvd_psnr.argtypes = [vp, vp, ci, ci, ci, ci, ci, vp, vp]
    lib.bsvd_ssim.argtypes = [vp, vp, ci, ci, ci, ci, ci, C.c_float, vp, vp]
    lib.bsvd_denoise_clip_u8.argtypes = [vp, vp, C.c_float, vp, ci, ci, ci, ci, vp]
    lib.bsvd_stream_push.argtypes = [vp, vp, vp, vp, ci, ci, ci, cip, vp]
    lib.bsvd_reset.argtypes = [vp]
    lib.bsvd_stream_graph_replays.argtypes = [vp]
    lib.bsvd_stream_graph_replays.restype = C.c_longlong
    lib.bsvd_last_launch_count.argtypes = [vp]
    lib.bsvd_workspace_bytes.argtypes = [vp]
    lib.bsvd_workspace_bytes.restype = C.c_size_t
    lib.bsvd_overflow_flag.argtypes = [vp, cip, ci]
    sz, cu = C.c_size_t, C.c_uint
    lib.bsvd_peer_create.argtypes = [ci, ci, sz, ci, C.POINTER(vp)]
    lib.bsvd_peer_get_handle.argtypes = [vp, vp]
    lib.bsvd_peer_open.argtypes = [vp, vp]
    lib.bsvd_peer_local_data.argtypes = [vp]
    lib.bsvd_peer_local_data.restype = vp
    lib.bsvd_peer_put.argtypes = [vp, ci, sz, vp, sz, vp]
    lib.bsvd_peer_put2d.argtypes = [vp, ci, sz, sz, vp, sz, sz, sz, vp]
    lib.bsvd_peer_put3d.argtypes = [vp, ci, sz, sz, sz, vp, sz, sz, sz, sz, sz, vp]
    lib.bsvd_peer_signal.argtypes = [vp, ci, ci, cu, vp]
    lib.bsvd_peer_wait.argtypes = [vp, ci, cu, vp]
    lib.bsvd_peer_read_flag.argtypes = [vp, ci, C.POINTER(cu)]
    lib.bsvd_peer_destroy.argtypes = [vp]
    lib.bsvd_set_profiling.argtypes = [vp, ci]
    lib.bsvd_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float), ci, cip]
    lib.bsvd_stage_info.argtypes = [vp, ci, cip, cip, cip, cip, cip]
    lib.bsvd_last_stage_ms.restype = C.c_float
    lib.bsvd_conv_stage.argtypes = [C.POINTER(BsvdConvDesc), vp, vp, vp, vp, vp, vp]
    if path is None:
        _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise BsvdError(load_library().bsvd_last_error().decode("utf-8", "replace"))
