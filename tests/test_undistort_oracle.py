"""The undistortion oracle against outputs of the reference's LensDistortionCorrector (cv2.undistortPoints), and the host-side
surface of calibration.lens_distortion (no GPU)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import undistort_oracle
from office_person_detection_vit_b200.calibration import CameraIntrinsics, DistortionParams, LensDistortionCorrector

G = np.load(Path(__file__).parent / "golden" / "undistort_golden.npz")


@pytest.mark.parametrize("i", range(len(G["cases"])))
def test_oracle_matches_cv2(i):
    got = undistort_oracle.undistort_points(G[f"pts{i}"], *G["cases"][i])
    # float64 on both sides; OpenCV's compiler may contract differently: a few ulps of a ~1e3 px coordinate
    np.testing.assert_allclose(got, G[f"und{i}"], rtol=0, atol=1e-9)


def test_params_surface():
    d = DistortionParams.from_array([0.1, 0.2, 0.3, 0.4, 0.5])
    assert (d.k1, d.k2, d.p1, d.p2, d.k3) == (0.1, 0.2, 0.3, 0.4, 0.5)
    assert list(d.to_array()) == [0.1, 0.2, 0.3, 0.4, 0.5]
    assert DistortionParams.from_array([0.1, 0.2]).to_dict() == {"k1": 0.1, "k2": 0.2, "k3": 0.0, "p1": 0.0, "p2": 0.0}
    assert DistortionParams().is_zero() and not d.is_zero()
    ci = CameraIntrinsics.from_config({"focal_length_x": 1000.0, "center_x": 600.0, "distortion": {"k1": -0.1}})
    assert ci.fx == 1000.0 and ci.fy == 1250.0 and ci.cx == 600.0 and ci.distortion.k1 == -0.1
    assert ci.get_camera_matrix().tolist() == [[1000.0, 0, 600.0], [0, 1250.0, 360.0], [0, 0, 1]]
    assert CameraIntrinsics.from_config({"distortion": [0.1, 0.2, 0.3, 0.4]}).distortion.p2 == 0.4


def test_disabled_corrector_is_identity_without_gpu():
    c = LensDistortionCorrector(CameraIntrinsics())
    assert not c.enabled
    assert c.undistort_point((3.0, 4.0)) == (3.0, 4.0)
    p = np.arange(12.0).reshape(6, 1, 2)
    assert c.undistort_points(p).shape == (6, 2)
