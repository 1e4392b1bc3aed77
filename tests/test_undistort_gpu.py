"""GPU undistortion (csrc/pwa.cu `undistort_points_kernel`) against the oracle and the cv2 golden, and PWA / TPS with a corrector."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(Path(__file__).parent / "golden" / "undistort_golden.npz")


def _corrector(c):
    from office_person_detection_vit_b200.calibration import CameraIntrinsics, DistortionParams, LensDistortionCorrector
    fx, fy, cx, cy, k1, k2, p1, p2, k3 = c
    return LensDistortionCorrector(CameraIntrinsics(fx=fx, fy=fy, cx=cx, cy=cy, distortion=DistortionParams(k1=k1, k2=k2, k3=k3, p1=p1, p2=p2)))


@pytest.mark.parametrize("i", range(len(G["cases"])))
def test_kernel_matches_cv2_and_oracle(i):
    from oracle import undistort_oracle
    corr = _corrector(G["cases"][i])
    got = corr.undistort_points(G[f"pts{i}"])
    np.testing.assert_allclose(got, G[f"und{i}"], rtol=0, atol=1e-9)     # the reference's cv2 output
    # same operation order as the oracle; the compiler may fuse a*b+c on the device: a few ulps
    np.testing.assert_allclose(got, undistort_oracle.undistort_points(G[f"pts{i}"], *G["cases"][i]), rtol=0, atol=1e-9)
    x = corr.undistort_point(tuple(G[f"pts{i}"][7]))
    assert x == (got[7, 0], got[7, 1])


def test_large_and_empty():
    import torch
    from oracle import undistort_oracle
    corr = _corrector(G["cases"][1])
    assert corr.undistort_tensor(torch.empty(0, 2, dtype=torch.float64, device="cuda")).shape == (0, 2)
    g = torch.Generator(device="cuda").manual_seed(5)
    pts = torch.rand(1_000_003, 2, generator=g, device="cuda", dtype=torch.float64) * torch.tensor([1280.0, 720.0], device="cuda", dtype=torch.float64)
    got = corr.undistort_tensor(pts).cpu().numpy()
    np.testing.assert_allclose(got, undistort_oracle.undistort_points(pts.cpu().numpy(), *G["cases"][1]), rtol=0, atol=1e-9)
    boxes = torch.cat([pts[:1000] - 5.0, torch.full((1000, 2), 10.0, device="cuda", dtype=torch.float64)], dim=1)
    foot = torch.stack([boxes[:, 0] + boxes[:, 2] / 2, boxes[:, 1] + boxes[:, 3]], dim=1)
    assert torch.equal(corr.undistort_tensor(boxes, is_bbox=True), corr.undistort_tensor(foot))


@pytest.mark.parametrize("kind", ["pwa", "tps"])
def test_transformers_apply_the_corrector_first(kind):
    import torch
    from office_person_detection_vit_b200.transform import PiecewiseAffineTransformer, ThinPlateSplineTransformer
    P = np.load(Path(__file__).parent / "golden" / "pwa_golden.npz")
    corr = _corrector(G["cases"][1])
    cls = PiecewiseAffineTransformer if kind == "pwa" else ThinPlateSplineTransformer
    plain, with_c = cls(P["src"], P["dst"]), cls(P["src"], P["dst"], distortion_corrector=corr)
    pts = torch.from_numpy(G["pts1"][:300]).cuda()
    a = with_c.transform_points(pts)
    b = plain.transform_points(corr.undistort_tensor(pts))
    assert torch.equal(a, b)
    info = with_c.get_info()
    assert info["distortion_correction_enabled"] and info["distortion_params"]["k1"] == G["cases"][1][4]
    assert not plain.get_info()["distortion_correction_enabled"]
    r = with_c.transform_pixel((float(pts[3, 0]), float(pts[3, 1])))
    np.testing.assert_allclose(r.floor_coords_px, a[3].cpu().numpy(), rtol=0, atol=1e-9)
