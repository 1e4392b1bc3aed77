"""Lens-distortion correction of points on the GPU, surface of the reference's src/calibration/lens_distortion.py:
`DistortionParams` (:23-77), `CameraIntrinsics` (:80-135), `LensDistortionCorrector.undistort_points / undistort_point`
(:138-203).  The reference calls cv2.undistortPoints(pts, K, dist, P=K); `undistort_points_kernel` (csrc/pwa.cu) restates
OpenCV's five fixed-point iterations in float64.  Image undistortion and the grid visualisation are not on the Phase 2 -> 3
path and are not built.  Used by PiecewiseAffineTransformer / ThinPlateSplineTransformer (`distortion_corrector=`)."""

from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field

import numpy as np

from .. import _lib

logger = logging.getLogger(__name__)

_P = C.c_void_p
_lib.register("opd_undistort_points_f64", C.c_int, [C.c_double] * 9 + [_P, C.c_int32, C.c_int64, _P, _P])


@dataclass
class DistortionParams:
    """OpenCV distortion model: radial k1, k2, k3 and tangential p1, p2 (lens_distortion.py:23-77)."""

    k1: float = 0.0
    k2: float = 0.0
    k3: float = 0.0
    p1: float = 0.0
    p2: float = 0.0

    def to_array(self) -> np.ndarray:
        return np.array([self.k1, self.k2, self.p1, self.p2, self.k3], dtype=np.float64)

    @classmethod
    def from_array(cls, arr) -> "DistortionParams":
        arr = np.array(arr).flatten()
        if len(arr) >= 5:
            return cls(k1=arr[0], k2=arr[1], p1=arr[2], p2=arr[3], k3=arr[4])
        if len(arr) >= 4:
            return cls(k1=arr[0], k2=arr[1], p1=arr[2], p2=arr[3])
        if len(arr) >= 2:
            return cls(k1=arr[0], k2=arr[1])
        return cls()

    def is_zero(self) -> bool:
        return all(abs(v) < 1e-10 for v in (self.k1, self.k2, self.k3, self.p1, self.p2))

    def to_dict(self) -> dict[str, float]:
        return {"k1": self.k1, "k2": self.k2, "k3": self.k3, "p1": self.p1, "p2": self.p2}


@dataclass
class CameraIntrinsics:
    """fx, fy, cx, cy [pixels], image size, distortion (lens_distortion.py:80-135)."""

    fx: float = 1250.0
    fy: float = 1250.0
    cx: float = 640.0
    cy: float = 360.0
    width: int = 1280
    height: int = 720
    distortion: DistortionParams = field(default_factory=DistortionParams)

    def get_camera_matrix(self) -> np.ndarray:
        return np.array([[self.fx, 0, self.cx], [0, self.fy, self.cy], [0, 0, 1]], dtype=np.float64)

    @classmethod
    def from_config(cls, config: dict) -> "CameraIntrinsics":
        dist_config = config.get("distortion", {})
        if isinstance(dist_config, list):
            distortion = DistortionParams.from_array(dist_config)
        elif isinstance(dist_config, dict):
            distortion = DistortionParams(k1=dist_config.get("k1", 0.0), k2=dist_config.get("k2", 0.0), k3=dist_config.get("k3", 0.0),
                                          p1=dist_config.get("p1", 0.0), p2=dist_config.get("p2", 0.0))
        else:
            distortion = DistortionParams()
        return cls(fx=config.get("focal_length_x", 1250.0), fy=config.get("focal_length_y", 1250.0), cx=config.get("center_x", 640.0),
                   cy=config.get("center_y", 360.0), width=config.get("image_width", 1280), height=config.get("image_height", 720),
                   distortion=distortion)


class LensDistortionCorrector:
    def __init__(self, intrinsics: CameraIntrinsics):
        self.intrinsics = intrinsics
        self.camera_matrix = intrinsics.get_camera_matrix()
        self.dist_coeffs = intrinsics.distortion.to_array()
        self.enabled = not intrinsics.distortion.is_zero()
        if self.enabled:
            logger.info(f"LensDistortionCorrector enabled: {intrinsics.distortion.to_dict()}")
        else:
            logger.info("LensDistortionCorrector disabled (zero distortion)")

    def undistort_tensor(self, points, *, is_bbox: bool = False):
        """[N,2] points (or [N,4] boxes: their foot points) float64 CUDA tensor -> [N,2] corrected pixel coordinates."""
        torch = _lib.require_cuda()
        cols = 4 if is_bbox else 2
        if points.dim() != 2 or points.shape[1] != cols or not points.is_cuda or points.dtype != torch.float64:
            raise ValueError(f"points must be a float64 CUDA tensor of shape [N,{cols}]")
        pts = points.contiguous()
        out = torch.empty((pts.shape[0], 2), dtype=torch.float64, device=pts.device)
        i, d = self.intrinsics, self.intrinsics.distortion
        k = (d.k1, d.k2, d.p1, d.p2, d.k3) if self.enabled else (0.0, 0.0, 0.0, 0.0, 0.0)
        with torch.cuda.device(pts.device):
            rc = _lib.lib().opd_undistort_points_f64(float(i.fx), float(i.fy), float(i.cx), float(i.cy), float(k[0]), float(k[1]), float(k[2]),
                                                     float(k[3]), float(k[4]), _lib.ptr(pts), int(is_bbox), pts.shape[0], _lib.ptr(out),
                                                     _lib.stream_ptr())
        _lib.check(rc, "opd_undistort_points_f64")
        return out

    def undistort_points(self, points: np.ndarray) -> np.ndarray:
        """(N, 2) or (N, 1, 2) -> (N, 2) (lens_distortion.py:156-184); the input is returned unchanged when disabled."""
        if not self.enabled:
            return points.reshape(-1, 2) if points.ndim == 3 else points
        torch = _lib.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 2)).to(torch.device("cuda", torch.cuda.current_device()))
        return self.undistort_tensor(t).cpu().numpy()

    def undistort_point(self, point: tuple[float, float]) -> tuple[float, float]:
        if not self.enabled:
            return point
        out = self.undistort_points(np.array([[point]], dtype=np.float64))
        return (float(out[0, 0]), float(out[0, 1]))
