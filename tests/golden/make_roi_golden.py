"""Generates tests/golden/roi_golden.npz with the reference's OWN FeatureExtractor.extract_roi_features
(src/tracking/feature_extractor.py:39-88) on a seeded feature map and boxes (inside, partly outside, degenerate and
sub-cell boxes).  Run here (CPU, needs /root/reference): python tests/golden/make_roi_golden.py"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from src.tracking.feature_extractor import FeatureExtractor  # noqa: E402


def main():
    rng = np.random.default_rng(11)
    feat = rng.standard_normal((13, 21, 64)).astype(np.float32)
    img = (720, 1280)
    boxes = [(100.0, 50.0, 200.0, 400.0), (0.0, 0.0, 1280.0, 720.0), (-40.5, -10.25, 90.0, 60.0), (1200.0, 600.0, 300.0, 300.0),
             (640.3, 360.7, 0.5, 0.5), (1279.9, 719.9, 10.0, 10.0), (300.0, 200.0, 0.0, 0.0), (31.0, 29.0, 29.9, 28.7)]
    boxes += [tuple(float(v) for v in (rng.uniform(-50, 1250), rng.uniform(-50, 700), rng.uniform(1, 400), rng.uniform(1, 500)))
              for _ in range(24)]
    out = FeatureExtractor().extract_roi_features(feat, boxes, img)
    np.savez_compressed(Path(__file__).resolve().parent / "roi_golden.npz", feat=feat, boxes=np.array(boxes, dtype=np.float64),
                        image_shape=np.array(img), features=out)
    print(out.shape, out.dtype)


if __name__ == "__main__":
    main()
