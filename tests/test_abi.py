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


def _exported_c_symbols(lib_path) -> set[str]:
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib_path)], capture_output=True, text=True, check=True).stdout
    return {ln.split()[-1] for ln in out.splitlines() if len(ln.split()) >= 3 and ln.split()[-2] in "TW" and
            ln.split()[-1].startswith("opd_")}


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    declared = _declared_symbols()
    assert declared, "no declarations parsed from include/opd_b200.h"
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # ... and nothing else: the C ABI of the product library is exactly the header (measurement probes live in libopd_probe.so)
    assert _exported_c_symbols(built_lib) == declared


def test_probe_library_is_separate(built_lib):
    probe = built_lib.with_name("libopd_probe.so")
    assert probe.exists()
    header = (built_lib.parent / "csrc" / "probe" / "opd_probe.h").read_text()
    declared = set(re.findall(r"\b(opd_\w+)\s*\(", header))
    assert declared == {"opd_debug_mma_probe", "opd_halo_conv3x3_test", "opd_debug_pipe_probe"}
    assert _exported_c_symbols(probe) == declared
    assert not (_exported_c_symbols(built_lib) & _exported_c_symbols(probe))


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
