"""ctypes binding of the C ABI declared in include/vocalie_b200.h.

The product path has NO CPU fallback: if ``libvocalie_b200.so`` is missing or fails to load,
every entry point raises ``BackendUnavailableError``.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

from .errors import BackendUnavailableError

_PKG = Path(__file__).resolve().parent
# VT_LIB_PATH: A/B timing sessions load an alternative build of the same sources (tools/ab_build.sh); never set in
# production, tests or bench
import os as _os
LIB_PATH = Path(_os.environ["VT_LIB_PATH"]) if _os.environ.get("VT_LIB_PATH") else _PKG / "libvocalie_b200.so"
_lock = threading.Lock()
_lib = None

VT_OK = 0
VT_OPERAND_FP16 = 0
VT_OPERAND_BF16 = 1
VT_OPERAND_FP32 = 2
POST_RESULT_STRIDE = 8
ABI_VERSION = 3


class PostParams(C.Structure):
    """struct vt_post_params (include/vocalie_b200.h)."""
    _fields_ = [
        ("sr", C.c_int32), ("trim", C.c_int32), ("silence_threshold", C.c_float),
        ("min_silence_frames", C.c_int32), ("snap_radius", C.c_int32),
        ("fade_in_frames", C.c_int32), ("fade_out_frames", C.c_int32), ("stitch", C.c_int32),
        ("gap_frames", C.c_int32), ("normalize", C.c_int32), ("clip", C.c_int32),
        ("target_peak", C.c_double), ("concat", C.c_int32), ("out_pcm16", C.c_int32),
        ("stitch_head", C.c_int32), ("stitch_tail", C.c_int32),
    ]


class HiftConfig(C.Structure):
    """struct vt_hift_config."""
    _fields_ = [("sampling_rate", C.c_int32), ("n_upsamples", C.c_int32), ("upsample_rates", C.c_int32 * 4),
                ("upsample_kernel_sizes", C.c_int32 * 4), ("source_resblock_kernel_sizes", C.c_int32 * 4), ("trim_fade", C.c_int32)]


class Tensor(C.Structure):
    """struct vt_tensor."""
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("shape", C.c_int64 * 4)]


_P = C.c_void_p
_I64 = C.c_int64
_SIGS = {
    "vt_abi_version": (C.c_int, []),
    "vt_last_error": (C.c_char_p, []),
    "vt_last_launch_count": (C.c_int, []),
    "vt_device_check": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "vt_post_workspace_bytes": (_I64, [C.c_int, _I64]),
    "vt_find_active_range": (C.c_int, [_P, _P, C.c_int, _I64, _I64, C.c_float, C.c_int, _P, _P, _I64, _P]),
    "vt_post_stats": (C.c_int, [_P, _P, C.c_int, _I64, _I64, C.c_float, _P, _P, _P, _I64, _P]),
    "vt_snap_zero_crossing": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, _P]),
    "vt_post_analyze": (C.c_int, [_P, _P, C.c_int, _I64, _I64, C.POINTER(PostParams), _P, _P, _I64, _P]),
    "vt_post_write": (C.c_int, [_P, _P, C.c_int, _I64, _I64, C.POINTER(PostParams), _P, _P, _I64, _P, _P, _P, _I64, _P]),
    "vt_post_process": (C.c_int, [_P, _P, C.c_int, _I64, _I64, C.POINTER(PostParams), _P, _I64, _P, _P, _P, _I64, _P]),
    "vt_wav_pcm16_header": (C.c_int, [_P, C.c_int, _P, _I64, _P]),
    "vt_pcm16_encode": (C.c_int, [_P, _P, _I64, _P]),
    "vt_pcm16_decode": (C.c_int, [_P, _P, _I64, _P]),
    "vt_hift_create": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(_P)]),
    "vt_hift_create_ex": (C.c_int, [C.POINTER(Tensor), C.c_int, C.c_int, C.POINTER(HiftConfig), C.POINTER(_P)]),
    "vt_hift_samples_per_frame": (C.c_int, [_P]),
    "vt_hift_sampling_rate": (C.c_int, [_P]),
    "vt_hift_destroy": (None, [_P]),
    "vt_hift_workspace_bytes": (_I64, [_P, C.c_int, _I64, _I64]),
    "vt_hift_forward": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.c_int, _P, _P, _P, C.c_uint64, _P, _P, _I64, _P]),
    "vt_resample": (C.c_int, [_P, _P, _P, C.c_int, _I64, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "vt_rms": (C.c_int, [_P, _P, C.c_int, _P, _P, _I64, _P]),
    "vt_hift_set_profiling": (C.c_int, [_P, C.c_int]),
    "vt_hift_read_profile": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                       C.POINTER(C.c_int)]),
    "vt_hift_read_timeline": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "vt_hift_read_tap": (_I64, [_P, C.c_char_p, C.c_int, _P, _I64, _P, _P]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def load_library():
    """Load the in-tree shared library (built by ``__graft_entry__.build()``)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise BackendUnavailableError(
                f"CUDA extension missing: {LIB_PATH} (run `python __graft_entry__.py` to build it); "
                "this path has no CPU fallback")
        try:
            lib = C.CDLL(str(LIB_PATH))
        except OSError as exc:
            raise BackendUnavailableError(f"cannot load {LIB_PATH}: {exc}") from exc
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise BackendUnavailableError(f"{LIB_PATH} does not export {name}") from exc
            fn.restype = res
            fn.argtypes = args
        if lib.vt_abi_version() != ABI_VERSION:
            raise BackendUnavailableError("ABI version mismatch between Python shim and libvocalie_b200.so")
        _lib = lib
        return lib


def check(rc: int, what: str = "") -> None:
    """Raise the boundary's single error type on a non-zero status."""
    if rc != VT_OK:
        msg = load_library().vt_last_error().decode(errors="replace")
        raise BackendUnavailableError(f"{what or 'vocalie_b200'} failed (status {rc}): {msg}")
