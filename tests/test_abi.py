"""CPU: libopd_b200.so builds for sm_100a, loads, and exports every symbol include/opd_b200.h declares.
No compute call is made here (no GPU)."""

from __future__ import annotations

import ctypes
import re
import subprocess

from .conftest import ROOT


def _declared_symbols() -> set[str]:
    text = (ROOT / "include" / "opd_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(opd_[a-z0-9_]+)\s*\(", text))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/opd_b200.h"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_header(built_lib):
    from office_person_detection_vit_b200 import _lib
    import office_person_detection_vit_b200.detection  # noqa: F401  (registers the detector entry points)
    import office_person_detection_vit_b200.transform  # noqa: F401  (homography / piecewise-affine / TPS / undistortion entry points)
    import office_person_detection_vit_b200.zone  # noqa: F401

    assert set(_lib.exported_symbols()) == _declared_symbols()
    assert _lib.lib().opd_version() == 1


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(built_lib)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_side_errors_without_gpu(built_lib):
    """Argument validation happens before any CUDA call."""
    from office_person_detection_vit_b200 import _lib

    out = ctypes.c_void_p()
    rc = _lib.lib().opd_zone_table_create(None, None, None, 65, 0, 0, ctypes.byref(out))
    assert rc == -1 and b"Z=65" in _lib.lib().opd_last_error()
