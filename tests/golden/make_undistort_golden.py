"""Generates tests/golden/undistort_golden.npz by running the REFERENCE's LensDistortionCorrector (cv2.undistortPoints) in the
build container: `python tests/golden/make_undistort_golden.py`.  Needs /root/reference and cv2; the tests only read the .npz."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from src.calibration.lens_distortion import CameraIntrinsics, DistortionParams, LensDistortionCorrector  # noqa: E402

rng = np.random.default_rng(20260118)
cases = [
    # fx, fy, cx, cy, k1, k2, p1, p2, k3
    (1250.0, 1250.0, 640.0, 360.0, -0.1, 0.0, 0.0, 0.0, 0.0),       # the config.yaml example values (k1 only)
    (1250.0, 1250.0, 640.0, 360.0, -0.28, 0.09, 0.001, -0.0005, -0.01),
    (900.0, 880.0, 655.5, 349.25, 0.15, -0.05, -0.002, 0.003, 0.004),  # pincushion + tangential
    (1250.0, 1250.0, 640.0, 360.0, -2.5, 0.0, 0.0, 0.0, 0.0),       # strong barrel: icdist < 0 at the image corners
]
out = {"cases": np.array(cases)}
for i, c in enumerate(cases):
    fx, fy, cx, cy, k1, k2, p1, p2, k3 = c
    corr = LensDistortionCorrector(CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, distortion=DistortionParams(k1=k1, k2=k2, k3=k3, p1=p1, p2=p2)))
    pts = np.concatenate([rng.uniform([-50, -50], [1330, 770], size=(500, 2)),
                          np.array([[cx, cy], [0.0, 0.0], [1279.0, 719.0], [640.0, 0.0], [0.0, 360.0]])])
    out[f"pts{i}"] = pts
    out[f"und{i}"] = corr.undistort_points(pts)
np.savez_compressed(Path(__file__).parent / "undistort_golden.npz", **out)
print({k: v.shape for k, v in out.items()})
