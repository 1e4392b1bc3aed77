"""Stage 1 of the phase golden (runs on a B200): this package's DetectionPhase on four synthetic 1280x720 camera frames ->
tests/golden/phase_detections.json, the detections the reference's own phases are then run on (make_phase_golden.py, stage 2,
runs in the CPU container where /root/reference is importable).

    python tests/golden/dump_phase_detections.py [out.json]
"""

from __future__ import annotations

import json
import logging
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

FRAMES = dict(n=4, h=720, w=1280, seed=7)
THRESHOLD = 0.5
TIMESTAMPS = ["2025/08/26 16:05:00", "2025/08/26 16:10:00", "2025/08/26 16:15:00", "2025/08/26 16:20:00"]


def main() -> None:
    from office_person_detection_vit_b200.detection import ViTDetector
    from office_person_detection_vit_b200.detection.synthetic import random_init_state_dict, synthetic_frames
    from office_person_detection_vit_b200.phases import DetectionPhase

    out = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "tests" / "golden" / "phase_detections.json"
    frames = synthetic_frames(FRAMES["n"], FRAMES["h"], FRAMES["w"], seed=FRAMES["seed"])
    det = ViTDetector(confidence_threshold=THRESHOLD, state_dict=random_init_state_dict(0))
    phase = DetectionPhase({"detection": {"confidence_threshold": THRESHOLD}}, logging.getLogger("dump"), detector=det)
    phase.initialize()
    results = phase.execute([(i * 150, TIMESTAMPS[i], frames[i]) for i in range(FRAMES["n"])])
    data = {"frames": FRAMES, "threshold": THRESHOLD, "weights": "random_init_state_dict(0)",
            "results": [{"frame_number": fn, "timestamp": ts,
                         "detections": [{"bbox": list(d.bbox), "confidence": d.confidence, "class_id": d.class_id,
                                         "class_name": d.class_name, "camera_coords": list(d.camera_coords)} for d in dets]}
                        for fn, ts, dets in results]}
    out.write_text(json.dumps(data, indent=0))
    print(f"wrote {out}: {[len(r['detections']) for r in data['results']]} detections")


if __name__ == "__main__":
    main()
