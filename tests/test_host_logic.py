"""CPU: host-side behaviour of the reference-shaped classes that needs no GPU — constructor validation
and error messages (reference tests/test_homography.py:85-95, tests/test_zone_classifier.py:28-32),
records, CSV layout (tests/test_aggregator.py:59-71)."""

from __future__ import annotations

import numpy as np
import pytest

from office_person_detection_vit_b200.aggregation import Aggregator
from office_person_detection_vit_b200.models import AggregationResult, Detection, FrameResult
from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer, TransformResult
from office_person_detection_vit_b200.zone import ZoneClassifier


def test_homography_rejects_bad_matrices():
    fm = FloorMapConfig()
    with pytest.raises(ValueError, match=r"3x3"):
        HomographyTransformer(np.eye(2), fm)
    with pytest.raises(ValueError, match=r"特異行列"):
        HomographyTransformer(np.zeros((3, 3)), fm)
    tr = HomographyTransformer(np.eye(3), fm)
    assert tr.get_info()["method"] == "homography" and tr.get_info()["floormap_size"] == (1878, 1369)
    assert tr._get_foot_point((100, 200, 50, 100)) == (125.0, 300.0)


def test_transform_result_defaults():
    r = TransformResult()
    assert r.floor_coords_px is None and r.floor_coords_mm is None
    assert r.is_valid is False and r.error_reason is None and r.is_within_bounds is False


def test_floormap_config_from_config():
    fm = FloorMapConfig.from_config({"image_width": 100, "image_height": 50, "image_x_mm_per_pixel": 10})
    assert (fm.width_px, fm.height_px, fm.scale_x_mm_per_px) == (100, 50, 10.0)
    assert fm.scale_x_m_per_px == pytest.approx(0.01) and fm.scale_x_px_per_m == pytest.approx(100.0)


@pytest.mark.parametrize("zones,pattern", [
    ({}, r".*リスト.*"),
    (["x"], r"辞書"),
    ([{"polygon": [[0, 0], [1, 0], [1, 1]]}], r"'id'"),
    ([{"id": "a"}], r"'polygon'"),
    ([{"id": "a", "polygon": [[0, 0], [1, 1]]}], r"3つの頂点"),
    ([{"id": "a", "polygon": [[0, 0], [1, 0], [1]]}], r"\[x, y\]形式"),
    ([{"id": "a", "polygon": [[0, 0], [1, 0], ["q", 1]]}], r"数値"),
    ([{"id": "a", "polygon": [[0, 0], [1, 0], [1, 1]]}, {"id": "a", "polygon": [[0, 0], [1, 0], [1, 1]]}], r"重複"),
    ([{"id": "a", "polygon": [[0, 0], [1, 0], [1, 1]], "priority": "high"}], r"priority"),
])
def test_zone_validation_errors(zones, pattern):
    with pytest.raises(ValueError, match=pattern):
        ZoneClassifier(zones)


def test_zone_metadata():
    zc = ZoneClassifier([{"id": "zone_a", "name": "A", "polygon": [(0, 0), (1, 0), (1, 1)], "priority": 2},
                         {"id": "zone_b", "polygon": [[0, 0], [2, 0], [2, 2]]}], allow_overlap=False)
    assert zc.get_all_zone_ids() == ["zone_a", "zone_b"] and zc.get_zone_count() == 2
    assert zc.get_zone_info("zone_b")["name"] == "zone_b" and zc.get_zone_info("zone_b")["priority"] is None
    assert zc.get_zone_info("nope") is None


def test_records():
    d = Detection(bbox=(1, 2, 3, 4), confidence=0.5, class_id=1, class_name="person", camera_coords=(2.5, 6))
    assert d.zone_ids == [] and d.floor_coords is None and d.track_id is None
    fr = FrameResult(frame_number=1, timestamp="12:00", detections=[d], zone_counts={})
    assert fr.detections[0] is d
    assert AggregationResult("12:00", "zone_a", 3).count == 3


def test_csv_layout(tmp_path):
    agg = Aggregator()
    agg.aggregate_histogram(["12:05", "12:00"], np.array([[1, 0, 2], [0, 3, 0]]), ["zone_a", "zone_b"])
    assert [(r.timestamp, r.zone_id, r.count) for r in agg.results] == [
        ("12:05", "zone_a", 1), ("12:05", "unclassified", 2), ("12:00", "zone_b", 3)]   # sparse storage
    out = tmp_path / "zone_counts.csv"
    agg.export_csv(str(out), zone_ids=["zone_a", "zone_b"])
    assert out.read_text().splitlines() == ["timestamp,zone_a,zone_b,unclassified", "12:00,0,3,0", "12:05,1,0,2"]


# ---- phases.py host logic (no GPU): config handling, the reference's error messages, the JSON writer ----------------------
def test_phase_config_errors_match_the_reference():
    import logging

    from office_person_detection_vit_b200.phases import DetectionPhase, TransformPhase, _cfg

    log = logging.getLogger("test_host_logic")
    assert _cfg({"a": {"b": 3}}, "a.b") == 3 and _cfg({"a": {"b": 3}}, "a.c", 7) == 7 and _cfg(None, "x", 1) == 1
    with pytest.raises(RuntimeError, match="initialize"):                      # detection.py:70-71
        DetectionPhase({}, log).execute([])
    with pytest.raises(ValueError, match="homography.matrix が設定されていません"):    # transform.py:144-145
        TransformPhase({"zones": []}, log).initialize()
    with pytest.raises(ValueError, match="3x3"):                               # transform.py:148-149
        TransformPhase({"homography": {"matrix": [[1, 0], [0, 1]]}}, log).initialize()
    with pytest.raises(RuntimeError, match="Not initialized"):                 # transform.py:266-267
        TransformPhase({}, log).execute([])


@pytest.mark.parametrize("name", ["grid16", "star16"])
def test_export_results_reproduces_the_reference_json(name, tmp_path):
    """TransformPhase.export_results on the records the REFERENCE's phases produced (tests/golden/phase_golden/<name>/
    frame_results.json, full precision) writes the reference's coordinate_transformations.json byte for byte
    (transform.py:398-531: key names, rounding, indent)."""
    import json
    import logging

    from office_person_detection_vit_b200.models import Detection, FrameResult
    from office_person_detection_vit_b200.phases import TransformPhase

    from .conftest import GOLDEN

    gold = GOLDEN / "phase_golden" / name
    c = json.loads((gold / "config.json").read_text())
    cfg = {"homography": {"matrix": c["homography"]}, "floormap": c["floormap"], "zones": c["zones"],
           "output": {"json_optimization": c["json_optimization"]}}
    p3 = TransformPhase(cfg, logging.getLogger("test_host_logic"))
    p3.initialize()                                                            # host only: validation, no device table yet
    frs = [FrameResult(frame_number=r["frame_number"], timestamp=r["timestamp"], zone_counts=r["zone_counts"],
                       detections=[Detection(bbox=tuple(d["bbox"]), confidence=d["confidence"], class_id=1, class_name="person",
                                             camera_coords=tuple(d["camera_coords"]), floor_coords=tuple(d["floor_coords"]),
                                             floor_coords_mm=tuple(d["floor_coords_mm"]), zone_ids=d["zone_ids"])
                                   for d in r["detections"]])
           for r in json.loads((gold / "frame_results.json").read_text())]
    p3.export_results(frs, tmp_path)
    assert (tmp_path / "coordinate_transformations.json").read_text(encoding="utf-8") == \
        (gold / "coordinate_transformations.json").read_text(encoding="utf-8")
