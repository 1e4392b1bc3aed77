"""CPU: the oracle (oracle/floor_oracle.c + its numpy twin) against outputs of the reference itself
(tests/golden/*.npz, produced by tests/golden/make_golden.py) and the reference's own KATs."""

from __future__ import annotations

import json

import numpy as np
import pytest

from oracle import floor_oracle as fo

from .conftest import GOLDEN


def test_transform_matches_reference(golden):
    _, g, _ = golden
    px, mm, within = fo.transform(g["H"], g["boxes"], is_bbox=True)
    # same float64 operations; BLAS may fuse the 3-term dot product, hence 1e-12 relative instead of ==
    np.testing.assert_allclose(px, g["floor_px"], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(mm, g["floor_mm"], rtol=1e-12, atol=1e-7)
    assert (within == g["within"]).all()


def test_transform_numpy_twin(golden):
    _, g, _ = golden
    np.testing.assert_allclose(fo.transform_np(g["H"], g["boxes"], True), g["floor_px"], rtol=1e-13, atol=1e-10)


def test_classify_matches_reference(golden):
    _, g, zones = golden
    idx, mask = fo.classify(g["floor_px"], zones)  # classify the reference's own floor points: bit exact
    assert (idx == g["zone_idx"]).all()
    assert (mask == g["zone_mask"]).all()


def test_point_in_polygon_twins_agree():
    rng = np.random.default_rng(7)
    zones = fo.star_zones(16, seed=2)
    pts = np.stack([rng.uniform(0, fo.MAP_W, 400), rng.uniform(0, fo.MAP_H, 400)], axis=1)
    for z in zones[:6]:
        poly = np.ascontiguousarray(np.array(z["polygon"], dtype=np.float64))
        for x, y in pts:
            c = fo.lib().oracle_point_in_polygon(float(x), float(y), poly.ctypes.data, len(poly))
            assert bool(c) == fo.point_in_polygon_py(float(x), float(y), [tuple(v) for v in poly])


def test_count_matches_reference(golden):
    _, g, zones = golden
    T = g["hist_single"].shape[0]
    assert (fo.count(zone_idx=g["zone_idx"], slot=g["frame_of"], Z=len(zones), T=T) == g["hist_single"]).all()
    assert (fo.count(zone_mask=g["zone_mask"], slot=g["frame_of"], Z=len(zones), T=T) == g["hist_multi"]).all()


def test_reference_kats():
    """tests/test_homography.py:97-154 and tests/test_zone_classifier.py:35-62 of the reference."""
    kat = json.loads((GOLDEN / "reference_kats.json").read_text())
    for key in ("identity_pixel", "scale_pixel"):
        px, _, _ = fo.transform(np.array(kat[key]["H"], dtype=np.float64), [kat[key]["point"]], is_bbox=False)
        assert px[0] == pytest.approx(kat[key]["floor_px"])
    px, mm, _ = fo.transform(np.eye(3), [kat["foot_point"]["bbox"]], is_bbox=True)
    assert px[0] == pytest.approx(kat["foot_point"]["foot"])
    _, mm, _ = fo.transform(np.eye(3), [kat["mm"]["point"]], is_bbox=False)
    assert mm[0] == pytest.approx(kat["mm"]["floor_mm"])
    zones = json.loads((GOLDEN / "kat_overlap.zones.json").read_text())
    idx, mask = fo.classify([[60.0, 60.0], [200.0, 200.0]], zones)
    assert mask[0] == 0b11 and idx[0] == 1  # both squares; priority 1 (zone_b) wins without overlap
    assert mask[1] == 0 and idx[1] == -1


def test_bounds_kats():
    """tests/test_homography.py:129-142,225-243: [0,w) x [0,h) half-open bounds."""
    pts = [[0.0, 0.0], [1877.9, 1368.9], [1878.0, 100.0], [100.0, 1369.0], [-0.1, 5.0]]
    _, _, within = fo.transform(np.eye(3), pts, is_bbox=False)
    assert within.tolist() == [1, 1, 0, 0, 0]


def test_min_edge_distance():
    zones = [{"id": "a", "polygon": [[0, 0], [10, 0], [10, 10], [0, 10]]}]
    d = fo.min_edge_distance([[5, 5], [5, 12], [13, 14]], zones)
    np.testing.assert_allclose(d, [5.0, 2.0, 5.0])


def test_fused_cpu_path_matches_parts():
    zones = fo.grid_zones(16)
    pts = fo.camera_points(2000, seed=3)
    idx, hist = fo.project_classify_count(fo.H_CONFIG, pts, zones)
    px, _, _ = fo.transform(fo.H_CONFIG, pts.astype(np.float64), is_bbox=False)
    idx2, _ = fo.classify(px, zones)
    assert (idx == idx2).all()
    assert (hist == fo.count(zone_idx=idx2, Z=16)[0]).all()
