"""ctypes wrapper of oracle/floor_oracle.c + the synthetic workload generators of SURVEY.md §8d.

TEST INFRASTRUCTURE — not imported by the product package.
"""

from __future__ import annotations

import ctypes as C
import math
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_SO = _DIR / "liboracle.so"
_lib = None

# scene constants and synthetic workload generators (data only, shared by both sides of every parity test)
from office_person_detection_vit_b200.scene import (  # noqa: E402,F401
    H_CONFIG, MAP_H, MAP_W, SX_MM, SY_MM, camera_points, grid_zones, star_zones)


def build() -> Path:
    src = _DIR / "floor_oracle.c"
    if not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR), "-s", "-B", "liboracle.so"], check=True)
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.oracle_point_in_polygon.restype = C.c_int
        _lib.oracle_point_in_polygon.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def pack_zones(zones: list[dict]):
    """zones (reference dict form) -> (verts [V,2] f64, offs [Z+1] i32, prio [Z] f64 with +inf for None)."""
    verts, offs = [], [0]
    for z in zones:
        verts.extend((float(x), float(y)) for x, y in z["polygon"])
        offs.append(len(verts))
    prio = [math.inf if z.get("priority") is None else float(z["priority"]) for z in zones]
    return (np.ascontiguousarray(np.array(verts, dtype=np.float64).reshape(-1, 2)),
            np.ascontiguousarray(np.array(offs, dtype=np.int32)),
            np.ascontiguousarray(np.array(prio, dtype=np.float64)))


def transform(H, rows, is_bbox: bool, width_px=MAP_W, height_px=MAP_H, sx=SX_MM, sy=SY_MM):
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    n = rows.shape[0]
    px = np.empty((n, 2)); mm = np.empty((n, 2)); within = np.empty((n,), dtype=np.uint8)
    Hc = np.ascontiguousarray(H, dtype=np.float64)
    lib().oracle_transform(_p(Hc), _p(rows), C.c_int64(n), C.c_int(int(is_bbox)), C.c_double(width_px),
                           C.c_double(height_px), C.c_double(sx), C.c_double(sy), _p(px), _p(mm), _p(within))
    return px, mm, within


def classify(pts, zones):
    """-> (zone_idx int32 [N] single-label answer, zone_mask uint64 [N] overlap answer)."""
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    verts, offs, prio = pack_zones(zones)
    n = pts.shape[0]
    idx = np.empty((n,), dtype=np.int32); mask = np.empty((n,), dtype=np.uint64)
    lib().oracle_classify(_p(pts), C.c_int64(n), _p(verts), _p(offs), _p(prio), C.c_int(len(zones)), _p(idx), _p(mask))
    return idx, mask


def count(zone_idx=None, zone_mask=None, slot=None, Z=0, T=1):
    n = len(zone_idx) if zone_idx is not None else len(zone_mask)
    hist = np.zeros((T, Z + 1), dtype=np.int64)
    zi = None if zone_idx is None else np.ascontiguousarray(zone_idx, dtype=np.int32)
    zm = None if zone_mask is None else np.ascontiguousarray(zone_mask, dtype=np.uint64)
    sl = None if slot is None else np.ascontiguousarray(slot, dtype=np.int32)
    lib().oracle_count(_p(zi), _p(zm), _p(sl), C.c_int64(n), C.c_int(Z), C.c_int(T), _p(hist))
    return hist


def min_edge_distance(pts, zones):
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    verts, offs, _ = pack_zones(zones)
    d = np.empty((pts.shape[0],))
    lib().oracle_min_edge_distance(_p(pts), C.c_int64(pts.shape[0]), _p(verts), _p(offs), C.c_int(len(zones)), _p(d))
    return d


def project_classify_count(H, pts_f32, zones, want_idx=True):
    """The fused path on float32 camera points, single thread (the CPU baseline the bench times)."""
    pts = np.ascontiguousarray(pts_f32, dtype=np.float32)
    verts, offs, prio = pack_zones(zones)
    n, Z = pts.shape[0], len(zones)
    idx = np.empty((n,), dtype=np.int32) if want_idx else None
    hist = np.zeros((Z + 1,), dtype=np.int64)
    Hc = np.ascontiguousarray(H, dtype=np.float64)
    lib().oracle_project_classify_count(_p(Hc), _p(pts), C.c_int64(n), _p(verts), _p(offs), _p(prio), C.c_int(Z),
                                        _p(idx), _p(hist))
    return idx, hist


# ----------------------------------------------------------------------------------------------------
# numpy restatement (second, independent statement of the same arithmetic; used to cross-check the C)
# ----------------------------------------------------------------------------------------------------
def point_in_polygon_py(x: float, y: float, polygon) -> bool:
    """zone_classifier.py:162-197, line by line, Python floats."""
    n = len(polygon)
    inside = False
    p1x, p1y = polygon[0]
    xinters = 0.0
    for i in range(1, n + 1):
        p2x, p2y = polygon[i % n]
        if y > min(p1y, p2y) and y <= max(p1y, p2y) and x <= max(p1x, p2x):
            if p1y != p2y:
                xinters = (y - p1y) * (p2x - p1x) / (p2y - p1y) + p1x
            if p1x == p2x or x <= xinters:
                inside = not inside
        p1x, p1y = p2x, p2y
    return inside


def transform_np(H, rows, is_bbox: bool):
    """homography.py:166-175 with NumPy (same calls as the reference)."""
    rows = np.asarray(rows, dtype=np.float64)
    foot = np.stack([rows[:, 0] + rows[:, 2] / 2, rows[:, 1] + rows[:, 3]], axis=1) if is_bbox else rows
    pts_h = np.hstack([foot, np.ones((len(foot), 1))])
    t = (np.asarray(H, dtype=np.float64) @ pts_h.T).T
    return t[:, :2] / t[:, 2:3]
