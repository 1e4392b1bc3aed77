"""Stage 2 of the phase golden (runs in the CPU container, /root/reference importable): the REFERENCE's own, unmodified
DetectionPhase -> TransformPhase -> AggregationPhase (src/pipeline/phases/detection.py:56-133, transform.py:257-330 + :398-531,
aggregation.py:26-91) on the detections this package's detector produced for four synthetic camera frames
(tests/golden/phase_detections.json, written on a B200 by dump_phase_detections.py).  The reference's DetectionPhase drives a stub
detector that replays those detections through `detect_with_features` (the reference's YOLO weights and `ultralytics` are not
available; SURVEY.md Appendix A stub plugin), Phase 3 / 4 run on the reference's own engines.

Writes, per zone configuration, tests/golden/phase_golden/<name>/{coordinate_transformations.json, zone_counts.csv,
frame_results.json}: the files tests/test_phases_gpu.py must reproduce with THIS package's phases and engines.

    python tests/golden/make_phase_golden.py
"""

from __future__ import annotations

import json
import logging
import shutil
import sys
import tempfile
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
GOLDEN = ROOT / "tests" / "golden"

for n in ["ultralytics", "matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.patches", "matplotlib.dates",
          "matplotlib.figure", "matplotlib.axes"]:
    sys.modules.setdefault(n, MagicMock())
sys.path.insert(0, str(REF))
sys.path.insert(1, str(ROOT))

from office_person_detection_vit_b200.scene import H_CONFIG, grid_zones, star_zones  # noqa: E402  (data only)

REAL_ZONES = [   # config.yaml:226-238 of the reference
    {"id": "zone_1", "name": "ゾーン1（左）", "polygon": [[859, 912], [1095, 912], [1095, 1350], [859, 1350]], "priority": 1},
    {"id": "zone_2", "name": "ゾーン2（中央）", "polygon": [[1095, 912], [1331, 912], [1331, 1350], [1095, 1350]], "priority": 2},
    {"id": "zone_3", "name": "ゾーン3（右）", "polygon": [[1331, 912], [1567, 912], [1567, 1350], [1331, 1350]], "priority": 3},
]
FLOORMAP = {"image_width": 1878, "image_height": 1369, "image_origin_x": 7, "image_origin_y": 9,
            "image_x_mm_per_pixel": 28.1926406926406, "image_y_mm_per_pixel": 28.241430700447}
GRID16 = [{**z, "name": z["id"]} for z in grid_zones(16)]   # the reference's Phase 4 log lines read zone["name"]
STAR16 = [{**z, "name": z["id"]} for z in star_zones(16, seed=2)]   # concave polygons, overlapping pairs, priority None
CONFIGS = {
    "real_zones": {"zones": REAL_ZONES, "json_optimization": {}},
    "star16": {"zones": STAR16, "json_optimization": {}},
    "grid16": {"zones": GRID16, "json_optimization": {}},
    "grid16_compact": {"zones": GRID16, "json_optimization": {"enabled": True, "coordinate_precision": 1, "compact_keys": True}},
}


def main() -> None:
    from src.config import ConfigManager
    from src.models import Detection
    from src.pipeline.phases import AggregationPhase, DetectionPhase, TransformPhase

    fixture = json.loads((GOLDEN / "phase_detections.json").read_text())
    logger = logging.getLogger("phase_golden")

    class ReplayDetector:
        """Stands in for YOLOv8Detector: replays the recorded detections of the frame it is handed (frames carry their index)."""

        def detect_with_features(self, frame):
            rec = fixture["results"][int(frame[0, 0, 0])]
            dets = [Detection(bbox=tuple(d["bbox"]), confidence=d["confidence"], class_id=d["class_id"], class_name=d["class_name"],
                              camera_coords=tuple(d["camera_coords"])) for d in rec["detections"]]
            return dets, np.zeros((len(dets), 256), np.float32)

    for name, cfg in CONFIGS.items():
        with tempfile.TemporaryDirectory() as tmp:
            tmp = Path(tmp)
            config = ConfigManager("nonexistent_config.yaml")
            config.set("homography", {"matrix": H_CONFIG.tolist()})
            config.set("floormap", FLOORMAP)
            config.set("zones", cfg["zones"])
            config.set("transform", {"method": "homography"})
            config.set("output", {"directory": str(tmp), "save_detection_images": False, "json_optimization": cfg["json_optimization"]})
            p2 = DetectionPhase(config, logger)
            p2.detector = ReplayDetector()
            p2.output_path = tmp
            frames = [(r["frame_number"], r["timestamp"], np.full((2, 2, 3), i, np.uint8)) for i, r in enumerate(fixture["results"])]
            detection_results = p2.execute(frames)
            p3 = TransformPhase(config, logger)
            p3.initialize()
            frame_results = p3.execute(detection_results)
            p3.export_results(frame_results, tmp)
            AggregationPhase(config, logger).execute(frame_results, tmp)
            out = GOLDEN / "phase_golden" / name
            out.mkdir(parents=True, exist_ok=True)
            shutil.copy(tmp / "coordinate_transformations.json", out / "coordinate_transformations.json")
            shutil.copy(tmp / "zone_counts.csv", out / "zone_counts.csv")
            if name in ("grid16", "star16"):   # full-precision records (the JSON rounds to 6 decimals)
                (out / "frame_results.json").write_text(json.dumps(
                    [{"frame_number": fr.frame_number, "timestamp": fr.timestamp, "zone_counts": fr.zone_counts,
                      "detections": [{"bbox": list(d.bbox), "confidence": d.confidence, "camera_coords": list(d.camera_coords),
                                      "floor_coords": list(d.floor_coords), "floor_coords_mm": list(d.floor_coords_mm),
                                      "zone_ids": d.zone_ids} for d in fr.detections]} for fr in frame_results]))
            (out / "config.json").write_text(json.dumps({"zones": cfg["zones"], "json_optimization": cfg["json_optimization"],
                                                         "homography": H_CONFIG.tolist(), "floormap": FLOORMAP}, ensure_ascii=False))
            print(name, [len(fr.detections) for fr in frame_results], [fr.zone_counts for fr in frame_results][:2])


if __name__ == "__main__":
    main()
