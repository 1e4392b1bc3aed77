"""Generates tests/golden/export_golden.json: the text the REFERENCE's TransformPhase.export_results writes
(src/pipeline/phases/transform.py:398-531) for synthetic detections, for the default and the compact-key options,
together with the arrays the exporter under test receives.  Run here (CPU): python tests/golden/make_export_golden.py"""

import json
import logging
import sys
import tempfile
from pathlib import Path
from unittest.mock import MagicMock

import numpy as np

for n in ["ultralytics", "matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.patches",
          "matplotlib.dates", "matplotlib.figure", "matplotlib.axes"]:
    sys.modules.setdefault(n, MagicMock())
sys.path.insert(0, "/root/reference")

from src.config.config_manager import ConfigManager  # noqa: E402
from src.models.data_models import Detection, FrameResult  # noqa: E402
from src.pipeline.phases.transform import TransformPhase  # noqa: E402
from src.transform.floormap_config import FloorMapConfig  # noqa: E402
from src.transform.homography import HomographyTransformer  # noqa: E402

H = [[-0.8795888447, -2.8974379541, 417.8510123786], [-1.5459702925, -3.4570021203, 1054.0107447082],
     [-0.0011928509, -0.0035480452, 1.0]]


def main():
    rng = np.random.default_rng(11)
    B, Q = 3, 6
    n_keep = np.array([4, 0, 6], dtype=np.int32)
    xywh = np.zeros((B, Q, 4)); score = np.zeros((B, Q)); foot = np.zeros((B, Q, 2))
    px = np.zeros((B, Q, 2)); mm = np.zeros((B, Q, 2)); zone_idx = np.full((B, Q), -1, dtype=np.int32)
    zone_ids = ["zone_a", "zone_b", "zone_c"]
    fm = FloorMapConfig(width_px=1878, height_px=1369, origin_x_px=7, origin_y_px=9, scale_x_mm_per_px=28.1926406926406,
                        scale_y_mm_per_px=28.241430700447)
    tr = HomographyTransformer(np.array(H), fm)
    frames = []
    for b in range(B):
        dets = []
        for r in range(n_keep[b]):
            x, y, w, h = rng.uniform(0, 1000), rng.uniform(0, 500), rng.uniform(10, 200), rng.uniform(20, 300)
            res = tr.transform_detection((x, y, w, h))
            z = int(rng.integers(-1, 3))
            xywh[b, r] = (x, y, w, h); score[b, r] = float(np.float32(rng.uniform(0.5, 1.0)))
            foot[b, r] = (x + w / 2, y + h); px[b, r] = res.floor_coords_px; mm[b, r] = res.floor_coords_mm; zone_idx[b, r] = z
            dets.append(Detection(bbox=(x, y, w, h), confidence=float(score[b, r]), class_id=1, class_name="person",
                                  camera_coords=(x + w / 2, y + h), floor_coords=res.floor_coords_px,
                                  floor_coords_mm=res.floor_coords_mm, zone_ids=[zone_ids[z]] if z >= 0 else []))
        frames.append(FrameResult(frame_number=100 + b, timestamp=f"2025/08/26 16:0{b}:00", detections=dets, zone_counts={}))
    out = {"inputs": {"n_keep": n_keep.tolist(), "xywh": xywh.tolist(), "score": score.tolist(), "foot": foot.tolist(),
                      "px": px.tolist(), "mm": mm.tolist(), "zone_idx": zone_idx.tolist(), "zone_ids": zone_ids,
                      "frame_numbers": [100, 101, 102], "timestamps": [f.timestamp for f in frames],
                      "transformer_info": json.loads(json.dumps(tr.get_info(), default=str))},
           "cases": []}
    for opt in ({}, {"enabled": True, "coordinate_precision": 1, "compact_keys": True, "exclude_px_coords": True},
                {"enabled": True, "coordinate_precision": 2, "compact_keys": False, "exclude_px_coords": False}):
        cfg = ConfigManager("nonexistent_config.yaml")
        cfg.set("output.json_optimization", opt)
        phase = TransformPhase(cfg, logging.getLogger("golden"))
        phase.transformer = tr
        phase.transform_method = "homography"
        with tempfile.TemporaryDirectory() as d:
            phase.export_results(frames, Path(d))
            text = (Path(d) / "coordinate_transformations.json").read_text(encoding="utf-8")
        out["cases"].append({"json_optimization": opt, "text": text})
    (Path(__file__).resolve().parent / "export_golden.json").write_text(json.dumps(out), encoding="utf-8")
    print("cases", len(out["cases"]), [len(c["text"]) for c in out["cases"]])


if __name__ == "__main__":
    main()
