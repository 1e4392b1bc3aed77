"""GPU: this package's Phase 2 / 3 / 4 (`office_person_detection_vit_b200.phases`) against
files written by the REFERENCE's own unmodified phases (tests/golden/make_phase_golden.py: DetectionPhase with a replay detector
-> TransformPhase -> export_results -> AggregationPhase, run in the CPU container on the detections recorded in
tests/golden/phase_detections.json).

  * on the recorded detections, `coordinate_transformations.json` and `zone_counts.csv` are reproduced BYTE FOR BYTE and the
    per-detection records (floor px / mm, zone ids, zone counts) to 1e-9 relative - through the engines' object surface
    (HomographyTransformer.transform_batch, ZoneClassifier.classify_batch, Aggregator.aggregate_frame / export_csv), not a
    re-typed loop;
  * the live chain (DetectionPhase on the same synthetic frames with the CUDA detector) reproduces the recorded detections
    (same counts, boxes within the bf16 envelope of DESIGN.md) and yields consistent Phase 3 / 4 results."""

from __future__ import annotations

import json
import logging

import numpy as np
import pytest

from .conftest import GOLDEN

pytestmark = pytest.mark.gpu

CONFIGS = ["real_zones", "star16", "grid16", "grid16_compact"]


def _config(name: str) -> dict:
    c = json.loads((GOLDEN / "phase_golden" / name / "config.json").read_text())
    return {"homography": {"matrix": c["homography"]}, "floormap": c["floormap"], "zones": c["zones"],
            "transform": {"method": "homography"}, "output": {"json_optimization": c["json_optimization"]}}


def _recorded():
    from office_person_detection_vit_b200.models import Detection

    fx = json.loads((GOLDEN / "phase_detections.json").read_text())
    results = [(r["frame_number"], r["timestamp"],
                [Detection(bbox=tuple(d["bbox"]), confidence=d["confidence"], class_id=d["class_id"], class_name=d["class_name"],
                           camera_coords=tuple(d["camera_coords"])) for d in r["detections"]]) for r in fx["results"]]
    return fx, results


@pytest.mark.parametrize("name", CONFIGS)
def test_phase3_phase4_reproduce_reference_files(name, built_lib, tmp_path):
    import torch

    from office_person_detection_vit_b200.phases import AggregationPhase, TransformPhase

    torch.cuda.init()
    cfg = _config(name)
    _, detection_results = _recorded()
    logger = logging.getLogger("test_phases")
    p3 = TransformPhase(cfg, logger)
    with pytest.raises(RuntimeError, match="Not initialized"):
        p3.execute(detection_results)
    p3.initialize()
    frame_results = p3.execute(detection_results)
    assert [fr.zone_counts for fr in frame_results] == [{}] * len(frame_results)     # Phase 3 leaves the counts to Phase 4
    p3.export_results(frame_results, tmp_path)
    agg = AggregationPhase(cfg, logger).execute(frame_results, tmp_path)
    gold = GOLDEN / "phase_golden" / name
    assert (tmp_path / "coordinate_transformations.json").read_text(encoding="utf-8") == \
        (gold / "coordinate_transformations.json").read_text(encoding="utf-8")
    assert (tmp_path / "zone_counts.csv").read_text(encoding="utf-8") == (gold / "zone_counts.csv").read_text(encoding="utf-8")
    assert len(agg.results) == sum(len(fr.zone_counts) for fr in frame_results)
    if (gold / "frame_results.json").exists():
        ref = json.loads((gold / "frame_results.json").read_text())
        assert len(ref) == len(frame_results)
        for fr, r in zip(frame_results, ref):
            assert (fr.frame_number, fr.timestamp) == (r["frame_number"], r["timestamp"])
            assert fr.zone_counts == r["zone_counts"] and list(fr.zone_counts) == list(r["zone_counts"])   # same key order too
            assert len(fr.detections) == len(r["detections"])
            for d, rd in zip(fr.detections, r["detections"]):
                assert d.zone_ids == rd["zone_ids"]
                np.testing.assert_allclose(d.floor_coords, rd["floor_coords"], rtol=1e-9, atol=1e-9)
                np.testing.assert_allclose(d.floor_coords_mm, rd["floor_coords_mm"], rtol=1e-9, atol=1e-7)
                np.testing.assert_allclose(d.camera_coords, rd["camera_coords"], rtol=0, atol=1e-9)


def test_live_chain_matches_recorded_detections(built_lib, tmp_path):
    """frames -> DetectionPhase (CUDA detector) -> TransformPhase -> AggregationPhase.  The recorded detections came from this
    very detector; kernels whose accumulation order changed since may move a box inside the bf16 envelope (DESIGN.md numerics)."""
    import torch

    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
    from office_person_detection_vit_b200.phases import AggregationPhase, DetectionPhase, TransformPhase

    torch.cuda.init()
    fx, recorded = _recorded()
    f = fx["frames"]
    frames = synthetic_frames(f["n"], f["h"], f["w"], seed=f["seed"])
    det = ViTDetector(confidence_threshold=fx["threshold"], state_dict=random_init_state_dict(0))
    logger = logging.getLogger("test_phases_live")
    p2 = DetectionPhase({"detection": {"confidence_threshold": fx["threshold"]}}, logger, detector=det)
    with pytest.raises(RuntimeError):
        DetectionPhase({}, logger).execute([])
    p2.initialize()
    sample = [(r[0], r[1], frames[i]) for i, r in enumerate(recorded)]
    sample.insert(2, (999, "2025/08/26 16:12:00", np.zeros((4, 4), np.uint8)))        # a broken frame: [] for it, the rest unaffected
    results = p2.execute(sample)
    assert results[2] == (999, "2025/08/26 16:12:00", [])
    results.pop(2)
    assert [(fn, ts) for fn, ts, _ in results] == [(fn, ts) for fn, ts, _ in recorded]
    n_live, n_rec = [len(d) for _, _, d in results], [len(d) for _, _, d in recorded]
    assert all(abs(a - b) <= max(3, 0.05 * b) for a, b in zip(n_live, n_rec)), (n_live, n_rec)
    if n_live == n_rec:
        worst = max(abs(a - b) for (_, _, dl), (_, _, dr) in zip(results, recorded) for x, y in zip(dl, dr)
                    for a, b in zip(x.bbox, y.bbox))
        assert worst < 20.0, worst
    cfg = _config("grid16")
    p3 = TransformPhase(cfg, logger)
    p3.initialize()
    frame_results = p3.execute(results)
    AggregationPhase(cfg, logger).execute(frame_results, tmp_path)
    for fr in frame_results:
        assert sum(fr.zone_counts.values()) == len(fr.detections)
        assert all(d.floor_coords is not None and d.floor_coords_mm is not None for d in fr.detections)
    rows = (tmp_path / "zone_counts.csv").read_text().splitlines()
    assert rows[0].split(",")[0] == "timestamp" and len(rows) == 1 + len(frame_results)
