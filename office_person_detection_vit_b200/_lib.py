"""ctypes binding of libopd_b200.so (C ABI: include/opd_b200.h).

Fails loudly when the library has not been built or cannot be loaded — there is no CPU fallback.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libopd_b200.so"

OPD_MAX_ZONES = 64


class OpdError(RuntimeError):
    """A libopd_b200 call returned a negative status."""


class FloorParams(C.Structure):
    _fields_ = [
        ("H", C.c_double * 9),
        ("scale_x_mm", C.c_double),
        ("scale_y_mm", C.c_double),
        ("map_w_px", C.c_double),
        ("map_h_px", C.c_double),
        ("input_is_bbox", C.c_int32),
        ("skip_projection", C.c_int32),
    ]


_lib = None

_P = C.c_void_p
_SIGNATURES = {
    "opd_version": (C.c_int, []),
    "opd_last_error": (C.c_char_p, []),
    "opd_launch_count": (C.c_int64, []),
    "opd_set_option": (C.c_int, [C.c_char_p, C.c_int32]),
    "opd_zone_table_create": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "opd_zone_table_destroy": (None, [_P]),
    "opd_zone_table_info": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]),
    "opd_floor_project_classify_count_f32": (C.c_int, [C.POINTER(FloorParams), _P, _P, _P, C.c_int64, C.c_int32,
                                                       _P, _P, _P, _P, _P, _P, _P]),
    "opd_floor_project_classify_count_f64": (C.c_int, [C.POINTER(FloorParams), _P, _P, _P, C.c_int64, C.c_int32,
                                                       _P, _P, _P, _P, _P, _P, _P]),
    "opd_zone_histogram": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P]),
    "opd_zone_combine_groups": (C.c_int, [_P, _P, C.c_int32, C.c_int64, _P, _P, _P]),
}


def exported_symbols() -> list[str]:
    """Every entry point include/opd_b200.h declares (used by the CPU-side load test)."""
    return sorted(_SIGNATURES) + sorted(_OPTIONAL)


_OPTIONAL: dict[str, tuple] = {}


def register(name: str, restype, argtypes) -> None:
    """Let sibling modules (detector engine) declare their entry points next to their wrappers."""
    _OPTIONAL[name] = (restype, argtypes)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.restype, fn.argtypes = restype, argtypes


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OpdError(
                f"{LIB_PATH} is missing: build it with `python -m office_person_detection_vit_b200.build` "
                "(there is no CPU fallback)"
            )
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in {**_SIGNATURES, **_OPTIONAL}.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = restype, argtypes
        _lib = handle
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().opd_last_error().decode("utf-8", "replace")
        raise OpdError(f"{what or 'libopd_b200'} failed ({rc}): {msg}")


def require_cuda():
    """torch is used for device memory and streams only."""
    import torch

    if not torch.cuda.is_available():
        raise OpdError("no CUDA device: office_person_detection_vit_b200 has no CPU path")
    return torch


def stream_ptr() -> int:
    import torch

    return int(torch.cuda.current_stream().cuda_stream)


def ptr(t) -> int | None:
    return None if t is None else int(t.data_ptr())
