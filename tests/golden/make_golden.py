"""Generate the golden fixtures of tests/golden/ by running THE REFERENCE ITSELF.

Run in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports the reference's own Python classes (src.transform.homography.HomographyTransformer,
src.zone.zone_classifier.ZoneClassifier, src.aggregation.aggregator.Aggregator) and stores their
outputs on seeded inputs.  The fixtures pin oracle/floor_oracle.c (tests/test_oracle.py) and are the
targets of the GPU parity tests; /root/reference is never read at test time.
"""

from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
sys.path.insert(0, str(OUT.parent.parent))

from src.aggregation.aggregator import Aggregator  # noqa: E402
from src.models.data_models import Detection  # noqa: E402
from src.transform.floormap_config import FloorMapConfig  # noqa: E402
from src.transform.homography import HomographyTransformer  # noqa: E402
from src.zone.zone_classifier import ZoneClassifier  # noqa: E402

from oracle.floor_oracle import H_CONFIG, grid_zones, star_zones  # noqa: E402

CONFIG_ZONES = [  # config.yaml:226-238
    {"id": "zone_1", "polygon": [[859, 912], [1095, 912], [1095, 1350], [859, 1350]], "priority": 1},
    {"id": "zone_2", "polygon": [[1095, 912], [1331, 912], [1331, 1350], [1095, 1350]], "priority": 2},
    {"id": "zone_3", "polygon": [[1331, 912], [1567, 912], [1567, 1350], [1331, 1350]], "priority": 3},
]
KAT_ZONES = [  # tests/test_zone_classifier.py:10-25
    {"id": "zone_a", "polygon": [[0, 0], [100, 0], [100, 100], [0, 100]], "priority": 2},
    {"id": "zone_b", "polygon": [[50, 50], [150, 50], [150, 150], [50, 150]], "priority": 1},
]


def run_case(name: str, H: np.ndarray, zones: list[dict], boxes: np.ndarray, frame_of: np.ndarray):
    fm = FloorMapConfig()
    tr = HomographyTransformer(H, fm)
    res = tr.transform_batch([tuple(map(float, b)) for b in boxes])
    px = np.array([r.floor_coords_px for r in res], dtype=np.float64)
    mm = np.array([r.floor_coords_mm for r in res], dtype=np.float64)
    within = np.array([r.is_within_bounds for r in res], dtype=np.uint8)
    ids = [z["id"] for z in zones]
    zc_single = ZoneClassifier(zones, allow_overlap=False)
    zc_multi = ZoneClassifier(zones, allow_overlap=True)
    single, multi = [], []
    for p in px:
        pt = (float(p[0]), float(p[1]))
        single.append(zc_single.classify(pt))
        multi.append(zc_multi.classify(pt))
    zone_idx = np.array([ids.index(s[0]) if s else -1 for s in single], dtype=np.int32)
    zone_mask = np.array([sum(1 << ids.index(z) for z in m) for m in multi], dtype=np.uint64)
    # counting through the reference's Aggregator, per frame
    T = int(frame_of.max()) + 1 if len(frame_of) else 1
    hist_single = np.zeros((T, len(zones) + 1), dtype=np.int64)
    hist_multi = np.zeros((T, len(zones) + 1), dtype=np.int64)
    for hist, answers in ((hist_single, single), (hist_multi, multi)):
        for t in range(T):
            dets = [Detection(bbox=tuple(boxes[i]), confidence=0.9, class_id=1, class_name="person",
                              camera_coords=(0.0, 0.0), zone_ids=list(answers[i]))
                    for i in np.nonzero(frame_of == t)[0]]
            for zid, c in Aggregator().get_zone_counts(dets).items():
                hist[t, len(zones) if zid == "unclassified" else ids.index(zid)] = c
    np.savez_compressed(OUT / f"{name}.npz", H=H, boxes=boxes, frame_of=frame_of.astype(np.int32), floor_px=px,
                        floor_mm=mm, within=within, zone_idx=zone_idx, zone_mask=zone_mask,
                        hist_single=hist_single, hist_multi=hist_multi)
    (OUT / f"{name}.zones.json").write_text(json.dumps(zones))
    print(name, len(boxes), "boxes;", int((zone_idx >= 0).sum()), "classified;", "hist row0", hist_single[0].tolist())


def random_boxes(rng, n):
    x = rng.uniform(-50, 1280, n); y = rng.uniform(-50, 720, n)
    w = rng.uniform(5, 300, n); h = rng.uniform(5, 400, n)
    return np.stack([x, y, w, h], axis=1)


def main():
    rng = np.random.default_rng(20251118)
    # 1. real-scene constants: config.yaml homography + 3 zones, 21-ish boxes per frame, 8 frames
    n = 170
    boxes = random_boxes(rng, n)
    run_case("scene_config", H_CONFIG, CONFIG_ZONES, boxes, rng.integers(0, 8, n))
    # 2. identity homography + the overlapping KAT squares; boxes whose foot points sweep the squares,
    #    including points exactly on edges and vertices (half-open rule)
    xs = np.array([0, 25, 50, 60, 75, 100, 125, 150, 151, 200, -1], dtype=np.float64)
    pts = np.array([(x, y) for x in xs for y in xs])
    boxes = np.stack([pts[:, 0] - 5, pts[:, 1] - 20, np.full(len(pts), 10.0), np.full(len(pts), 20.0)], axis=1)
    run_case("kat_overlap", np.eye(3), KAT_ZONES, boxes, np.zeros(len(boxes), dtype=np.int64))
    # 3. config homography + 16 / 64 synthetic polygons (concave, rotated, overlapping, None priorities)
    for Z, seed in ((16, 2), (64, 4)):
        n = 3000
        # foot points drawn so that most land on the floormap: sample floor px, pull back through H^-1
        fp = np.stack([rng.uniform(-100, 1978, n), rng.uniform(-100, 1469, n), np.ones(n)], axis=1)
        cam = (np.linalg.inv(H_CONFIG) @ fp.T).T
        cam = cam[:, :2] / cam[:, 2:3]
        w = rng.uniform(5, 200, n); h = rng.uniform(5, 300, n)
        boxes = np.stack([cam[:, 0] - w / 2, cam[:, 1] - h, w, h], axis=1)
        run_case(f"star{Z}", H_CONFIG, star_zones(Z, seed), boxes, rng.integers(0, 64, n))
    run_case("grid4", H_CONFIG, grid_zones(4), random_boxes(rng, 500), rng.integers(0, 4, 500))
    # 4. the reference's own known-answer vectors (tests/test_homography.py:97-154,176-193)
    kat = {
        "identity_pixel": {"H": np.eye(3).tolist(), "point": [100.0, 200.0], "floor_px": [100.0, 200.0]},
        "scale_pixel": {"H": [[2, 0, 100], [0, 2, 50], [0, 0, 1]], "point": [50.0, 100.0], "floor_px": [200.0, 250.0]},
        "foot_point": {"bbox": [100, 200, 50, 100], "foot": [125.0, 300.0]},
        "mm": {"point": [100.0, 100.0], "floor_mm": [100.0 * 28.1926406926406, 100.0 * 28.241430700447]},
    }
    (OUT / "reference_kats.json").write_text(json.dumps(kat, indent=1))


if __name__ == "__main__":
    main()
