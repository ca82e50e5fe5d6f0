"""ctypes binding of libhippo_b200.so (the C ABI declared in include/hippo_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, or
the current device is not a B200-class (sm_100) GPU, the calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libhippo_b200.so"

HIPPO_OK = 0
HIPPO_E_BADARG = -1
HIPPO_E_ARCH = -2
HIPPO_E_CUDA = -3
HIPPO_E_NCCL = -4
HIPPO_E_WORKSPACE = -5

HIPPO_F32, HIPPO_F64, HIPPO_BF16, HIPPO_I16 = 0, 1, 2, 3
HIPPO_TOPK_MAX = 16
ABI_VERSION = 2


class HippoError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"libhippo_b200 status {status}: {message}")
        self.status = status


class HippoBadArgument(HippoError, ValueError):
    pass


class HippoArchError(HippoError):
    pass


class StreamDesc(C.Structure):
    """hippo_stream_desc"""

    _fields_ = [
        ("ssim", C.c_void_p),
        ("frame_times", C.c_void_p),
        ("nframes", C.c_int64),
        ("pcm", C.c_void_p),
        ("e16", C.c_void_p),
        ("e512", C.c_void_p),
        ("ns", C.c_int64),
        ("nch", C.c_int32),
        ("pcm_dtype", C.c_int32),
        ("sample_rate", C.c_double),
        ("out_bounds", C.c_void_p),
        ("out_count", C.c_void_p),
        ("max_segments", C.c_int32),
        ("reserved", C.c_int32),
    ]


class SegmentState(C.Structure):
    """hippo_segment_state"""

    _fields_ = [("current_start", C.c_double), ("hint", C.c_int64), ("count", C.c_int32), ("done", C.c_int32)]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_SZ = C.c_size_t
_F32 = C.c_float
_F64 = C.c_double

# name -> (restype, argtypes); exactly the declarations of include/hippo_b200.h
SIGNATURES = {
    "hippo_abi_version": (_I32, []),
    "hippo_last_error": (C.c_char_p, []),
    "hippo_device_check": (_I32, []),
    "hippo_sm_count": (_I32, []),
    "hippo_bank_build": (_I32, [_P, _I32, _I64, _I32, _I64, _P, _P, _P, _P]),
    "hippo_topk_single_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "hippo_topk_single": (_I32, [_P, _P, _I64, _I32, _P, _I32, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "hippo_topk_batched_workspace_bytes": (_SZ, [_I64, _I32, _I32, _I32]),
    "hippo_topk_batched": (_I32, [_P, _P, _I64, _I32, _P, _I32, _I32, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "hippo_topk_rows_workspace_bytes": (_SZ, [_I64, _I32, _I32]),
    "hippo_topk_rows": (_I32, [_P, _I32, _I64, _I32, _I64, _P, _I32, _I32, _I64, _P, _P, _P, _P, _P, _SZ, _P]),
    "hippo_rescore": (_I32, [_P, _I32, _I64, _I32, _I64, _I64, _P, _I32, _I64, _I32, _P, _I32, _P, _P, _P]),
    "hippo_topk_merge": (_I32, [_P, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "hippo_topk_exchange_bytes": (_SZ, [_I32, _I32, _I32]),
    "hippo_topk_exchange_merge": (_I32, [_P, _I32, _I32, _I32, _P, _SZ, _I32, _I32, C.c_uint32, _P, _P, _P, _P]),
    "hippo_topk_batched_sharded": (_I32, [_P, _P, _I64, _I32, _P, _I32, _I32, _I64, _P, _P, _SZ, _I32, _I32, C.c_uint32,
                                          _P, _P, _P, _P, _SZ, _P]),
    "hippo_topk_single_sharded": (_I32, [_P, _P, _I64, _I32, _P, _I32, _I64, _P, _P, _SZ, _I32, _I32, C.c_uint32,
                                         _P, _P, _P, _P, _SZ, _P]),
    "hippo_scores_single": (_I32, [_P, _P, _I64, _I32, _P, _P, _P]),
    "hippo_topk_segmented_workspace_bytes": (_SZ, [_I64]),
    "hippo_topk_segmented": (_I32, [_P, _P, _I64, _I32, _P, _P, _I32, _I32, _P, _P, _P, _P, _SZ, _P]),
    "hippo_recall_windows": (_I32, [_P, _P, _I32, _I32, _P, _P, _P, _F64, _I32, _P, _P, _P, _P, _P, _P]),
    "hippo_consolidate_workspace_bytes": (_SZ, [_I64, _I32]),
    "hippo_consolidate": (_I32, [_P, _I64, _I32, _F32, _F32, _F32, _P, _P, _P, _P, _SZ, _P]),
    "hippo_consolidate_ex_workspace_bytes": (_SZ, [_I64, _I32, _I32, _I64]),
    "hippo_consolidate_ex": (_I32, [_P, _I64, _I32, _F32, _F32, _F32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "hippo_debug_consolidate_timing": (None, [_P]),
    "hippo_frame_pairs_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "hippo_frame_pairs": (_I32, [_P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _P, _SZ, _P]),
    "hippo_audio_energy": (_I32, [_P, _I32, _I64, _I32, _P, _P, _P]),
    "hippo_audio_levels": (_I32, [_P, _I32, _I64, _I32, _P, _P, _P, _P, _I32, _P, _P]),
    "hippo_segment_boundaries": (_I32, [_P, _I32, _F64, _F64, _F64, _F64, _P]),
    "hippo_pattern_separation_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "hippo_pattern_separation": (_I32, [_P, _I32, _I32, _I32, _I32, _P, _P, _I32, _I64, _I32, _F64, _F64, _F64, _F64, _F64,
                                        _I32, _P, _P, _P, _P, _P, _P, _I32, _P, _SZ, _P, _P]),
    "hippo_segment_boundaries_resume": (_I32, [_P, _I32, _P, _I64, _I32, _F64, _F64, _F64, _F64, _P]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every entry point. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m hippomm_b200.build` "
            "(nvcc, sm_100a). hippomm_b200 has no CPU fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.hippo_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.hippo_abi_version()} != {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status == HIPPO_OK:
        return
    msg = load().hippo_last_error().decode("utf-8", "replace")
    if status == HIPPO_E_BADARG:
        raise HippoBadArgument(status, msg)
    if status == HIPPO_E_ARCH:
        raise HippoArchError(status, msg)
    raise HippoError(status, msg)
