"""Piecewise-affine camera pixel -> floormap transform on the GPU (the reference's shipped default, config.yaml:91).

Keeps the surface of the reference's PiecewiseAffineTransformer (src/transform/piecewise_affine.py:51-346): constructor
`(src_points, dst_points, floormap_config, distortion_corrector)` with its ValueErrors (:80-84), `transform_pixel`,
`transform_detection`, `transform_batch`, `evaluate_training_error`, `get_info`, `save` / `load`,
`from_correspondence_file`, and the `PWATransformResult` record (:28-48).  The triangulation (scipy.spatial.Delaunay) and the
per-triangle affine matrices (numpy.linalg.lstsq) are built on the host exactly as the reference builds them (:86-125);
point location, the nearest-centroid extrapolation (:138-153), the affine map, bounds check and mm scale (:155-205) run in
`pwa_transform_kernel` (csrc/pwa.cu) in float64.

`transform_points` is the tensor entry: [N,2] float64 CUDA points (or [N,4] boxes) in, [N,2] floor pixels out; its output feeds
`ZoneClassifier.count(points, transformer=None)` directly, so the PWA variant of the Phase 2 -> 3 path stays on the device.
A `distortion_corrector` (calibration.LensDistortionCorrector; disabled in the shipped config) is applied on the device before the
transform, as the reference applies it per point (:163-167)."""

from __future__ import annotations

import ctypes as C
import json
import logging
import pickle
from dataclasses import dataclass
from pathlib import Path
from typing import TYPE_CHECKING, Sequence

import numpy as np

from .. import _lib
from ..calibration import lens_distortion as _lens_distortion  # noqa: F401  (registers opd_undistort_points_f64)

if TYPE_CHECKING:
    from .floormap_config import FloorMapConfig

logger = logging.getLogger(__name__)

_P = C.c_void_p
_lib.register("opd_pwa_table_create", C.c_int, [_P, _P, _P, C.c_int32, C.c_double, C.c_int32, C.POINTER(_P)])
_lib.register("opd_pwa_table_destroy", None, [_P])
_lib.register("opd_pwa_transform_f64", C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                                _P, _P, _P, _P, _P, _P])

_lib.register("opd_tps_table_create", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.POINTER(_P)])
_lib.register("opd_tps_table_destroy", None, [_P])
_lib.register("opd_tps_transform_f64", C.c_int, [_P, _P, C.c_int32, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P,
                                                _P])

FIND_SIMPLEX_EPS = 100.0 * float(np.finfo(np.float64).eps)   # scipy.spatial.Delaunay.find_simplex default tolerance


@dataclass
class PWATransformResult:
    """Result of transforming one point (piecewise_affine.py:28-48)."""

    floor_coords_px: tuple[float, float] | None = None
    floor_coords_mm: tuple[float, float] | None = None
    is_valid: bool = False
    error_reason: str | None = None
    is_within_bounds: bool = False
    triangle_index: int = -1
    is_extrapolated: bool = False


class PiecewiseAffineTransformer:
    """Delaunay triangulation of the correspondences + one affine map per triangle, evaluated on the GPU."""

    def __init__(self, src_points, dst_points, floormap_config: "FloorMapConfig | None" = None, distortion_corrector=None):
        from scipy.spatial import Delaunay

        self.src_points = np.array(src_points, dtype=np.float64)
        self.dst_points = np.array(dst_points, dtype=np.float64)
        self.floormap_config = floormap_config
        self.distortion_corrector = distortion_corrector
        if len(self.src_points) < 3:
            raise ValueError("最低3点の対応点が必要です")
        if len(self.src_points) != len(self.dst_points):
            raise ValueError("src_points と dst_points の数が一致しません")
        self.delaunay = Delaunay(self.src_points)
        self.affine_matrices = self._compute_affine_matrices()
        self._centroids = np.mean(self.src_points[self.delaunay.simplices], axis=1)
        self._handles: dict[int, int] = {}
        logger.info(f"PiecewiseAffineTransformer initialized: {len(self.src_points)} points, {len(self.delaunay.simplices)} triangles")

    def _compute_affine_matrices(self) -> list[np.ndarray]:
        """One 3x3 affine matrix per triangle, least squares on the augmented vertices (piecewise_affine.py:102-125)."""
        mats = []
        for simplex in self.delaunay.simplices:
            src_aug = np.vstack([self.src_points[simplex].T, np.ones(3)])
            dst_aug = np.vstack([self.dst_points[simplex].T, np.ones(3)])
            A, _, _, _ = np.linalg.lstsq(src_aug.T, dst_aug.T, rcond=None)
            mats.append(A.T)
        return mats

    # -- device table -----------------------------------------------------------------------------------------
    def _handle(self, device_index: int) -> int:
        h = self._handles.get(device_index)
        if h is None:
            bary = np.ascontiguousarray(self.delaunay.transform, dtype=np.float64)                       # [T,3,2]
            aff = np.ascontiguousarray(np.array(self.affine_matrices)[:, :2, :], dtype=np.float64)      # [T,2,3]
            cen = np.ascontiguousarray(self._centroids, dtype=np.float64)                                # [T,2]
            out = _P()
            _lib.check(_lib.lib().opd_pwa_table_create(bary.ctypes.data, aff.ctypes.data, cen.ctypes.data, len(aff), FIND_SIMPLEX_EPS,
                                                       device_index, C.byref(out)), "opd_pwa_table_create")
            h = self._handles[device_index] = out.value
        return h

    def __del__(self):
        try:
            for h in getattr(self, "_handles", {}).values():
                _lib.lib().opd_pwa_table_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    # -- tensor entry -----------------------------------------------------------------------------------------
    def transform_points(self, points, *, is_bbox: bool = False, with_mm: bool = False, with_bounds: bool = False,
                         with_triangles: bool = False):
        """[N,2] points (or [N,4] boxes with is_bbox) float64 CUDA tensor -> [N,2] floor pixels; optionally also floor mm, the
        in-bounds flag and (triangle index int32, extrapolated uint8): (px[, mm][, within][, tri, extrapolated])."""
        torch = _lib.require_cuda()
        cols = 4 if is_bbox else 2
        if points.dim() != 2 or points.shape[1] != cols or not points.is_cuda or points.dtype != torch.float64:
            raise ValueError(f"points must be a float64 CUDA tensor of shape [N,{cols}]")
        pts = points.contiguous()
        if self.distortion_corrector is not None and self.distortion_corrector.enabled:
            pts, is_bbox = self.distortion_corrector.undistort_tensor(pts, is_bbox=is_bbox), False   # :163-167, foot point first
        n, dev = pts.shape[0], pts.device
        px = torch.empty((n, 2), dtype=torch.float64, device=dev)
        mm = torch.empty((n, 2), dtype=torch.float64, device=dev) if with_mm else None
        inb = torch.empty((n,), dtype=torch.uint8, device=dev) if with_bounds else None
        tri = torch.empty((n,), dtype=torch.int32, device=dev) if with_triangles else None
        ext = torch.empty((n,), dtype=torch.uint8, device=dev) if with_triangles else None
        fm = self.floormap_config
        sx, sy = (float(fm.scale_x_mm_per_px), float(fm.scale_y_mm_per_px)) if fm else (1.0, 1.0)
        mw, mh = (float(fm.width_px), float(fm.height_px)) if fm else (float("inf"), float("inf"))
        with torch.cuda.device(dev):
            rc = _lib.lib().opd_pwa_transform_f64(self._handle(dev.index), _lib.ptr(pts), int(is_bbox), n, sx, sy, mw, mh, _lib.ptr(px),
                                                  _lib.ptr(mm), _lib.ptr(inb), _lib.ptr(tri), _lib.ptr(ext), _lib.stream_ptr())
        _lib.check(rc, "opd_pwa_transform_f64")
        out = [px]
        if with_mm:
            out.append(mm)
        if with_bounds:
            out.append(inb)
        if with_triangles:
            out += [tri, ext]
        return out[0] if len(out) == 1 else tuple(out)

    # -- reference surface ------------------------------------------------------------------------------------
    def _run(self, rows: np.ndarray, is_bbox: bool) -> list[PWATransformResult]:
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(dev)
        px, mm, inb, tri, ext = self.transform_points(t, is_bbox=is_bbox, with_mm=True, with_bounds=True, with_triangles=True)
        px, mm, inb, tri, ext = (a.cpu().numpy() for a in (px, mm, inb, tri, ext))
        has_fm = self.floormap_config is not None
        return [PWATransformResult(floor_coords_px=(float(px[i, 0]), float(px[i, 1])),
                                   floor_coords_mm=(float(mm[i, 0]), float(mm[i, 1])) if has_fm else None,
                                   is_valid=True, is_within_bounds=bool(inb[i]) if has_fm else True,
                                   triangle_index=int(tri[i]), is_extrapolated=bool(ext[i])) for i in range(len(rows))]

    def transform_pixel(self, image_point: tuple[float, float]) -> PWATransformResult:
        return self._run(np.array([[image_point[0], image_point[1]]], dtype=np.float64), False)[0]

    def transform_detection(self, bbox: tuple[float, float, float, float]) -> PWATransformResult:
        return self._run(np.array([bbox], dtype=np.float64), True)[0]

    def transform_batch(self, bboxes: Sequence[tuple[float, float, float, float]]) -> list[PWATransformResult]:
        """One launch for all boxes (the reference loops transform_detection, piecewise_affine.py:224-236)."""
        if len(bboxes) == 0:
            return []
        return self._run(np.array(bboxes, dtype=np.float64).reshape(-1, 4), True)

    def evaluate_training_error(self) -> dict:
        """Error on the correspondences themselves (piecewise_affine.py:238-262), one launch."""
        res = self._run(self.src_points, False)
        errors = np.array([np.sqrt((r.floor_coords_px[0] - d[0]) ** 2 + (r.floor_coords_px[1] - d[1]) ** 2)
                           for r, d in zip(res, self.dst_points)])
        if errors.size == 0:
            return {"rmse": 0.0, "max_error": 0.0, "mean_error": 0.0}
        return {"rmse": float(np.sqrt(np.mean(errors ** 2))), "max_error": float(np.max(errors)), "mean_error": float(np.mean(errors)),
                "min_error": float(np.min(errors)), "std_error": float(np.std(errors)), "num_points": len(errors)}

    def get_info(self) -> dict:
        return {"method": "piecewise_affine", "num_points": len(self.src_points), "num_triangles": len(self.delaunay.simplices),
                "training_error": self.evaluate_training_error(), "distortion_correction_enabled": self.distortion_corrector is not None,
                **({"distortion_params": self.distortion_corrector.intrinsics.distortion.to_dict()} if self.distortion_corrector is not None else {})}

    def save(self, path: Path | str) -> None:
        with open(path, "wb") as f:
            pickle.dump({"src_points": self.src_points, "dst_points": self.dst_points}, f)
        logger.info(f"PWA model saved to {path}")

    @classmethod
    def load(cls, path: Path | str, floormap_config=None, distortion_corrector=None) -> "PiecewiseAffineTransformer":
        with open(path, "rb") as f:
            data = pickle.load(f)
        return cls(data["src_points"], data["dst_points"], floormap_config, distortion_corrector)

    @classmethod
    def from_correspondence_file(cls, file_path: Path | str, floormap_config=None, distortion_corrector=None) -> "PiecewiseAffineTransformer":
        with open(file_path, encoding="utf-8") as f:
            data = json.load(f)
        points = data.get("point_correspondences", [])
        return cls(np.array([p["src_point"] for p in points]), np.array([p["dst_point"] for p in points]), floormap_config,
                   distortion_corrector)


class ThinPlateSplineTransformer:
    """Thin-plate-spline camera -> floormap transform on the GPU, surface of the reference's class
    (src/transform/piecewise_affine.py:398-590): constructor `(src_points, dst_points, floormap_config, regularization,
    distortion_corrector)`, `transform_pixel / transform_detection / transform_batch`, `evaluate_training_error`, `get_info`,
    `from_correspondence_file`.  The coefficients are solved on the host exactly as the reference solves them (:445-485); the
    per-point radial-basis sum runs in `tps_transform_kernel` (csrc/pwa.cu), float64, in the reference's summation order."""

    def __init__(self, src_points, dst_points, floormap_config: "FloorMapConfig | None" = None, regularization: float = 0.0,
                 distortion_corrector=None):
        self.src_points = np.array(src_points, dtype=np.float64)
        self.dst_points = np.array(dst_points, dtype=np.float64)
        self.floormap_config = floormap_config
        self.regularization = regularization
        self.distortion_corrector = distortion_corrector
        if len(self.src_points) < 3:
            raise ValueError("最低3点の対応点が必要です")
        self.weights_x, self.weights_y, self.affine_x, self.affine_y = self._compute_tps_coefficients()
        self._handles: dict[int, int] = {}
        logger.info(f"ThinPlateSplineTransformer initialized with {len(self.src_points)} points")

    @staticmethod
    def _radial_basis(r: np.ndarray) -> np.ndarray:
        mask = r > 0
        result = np.zeros_like(r)
        result[mask] = r[mask] ** 2 * np.log(r[mask])
        return result

    def _compute_tps_coefficients(self):
        """[K + lambda I, P; P^T, 0] [w; a] = [v; 0], one solve per output coordinate (piecewise_affine.py:445-485)."""
        n = len(self.src_points)
        K = np.zeros((n, n))
        for i in range(n):
            for j in range(n):
                if i != j:
                    r = np.linalg.norm(self.src_points[i] - self.src_points[j])
                    K[i, j] = self._radial_basis(np.array([r]))[0]
        P = np.hstack([np.ones((n, 1)), self.src_points])
        L = np.zeros((n + 3, n + 3))
        L[:n, :n] = K + self.regularization * np.eye(n)
        L[:n, n:] = P
        L[n:, :n] = P.T
        v_x, v_y = np.zeros(n + 3), np.zeros(n + 3)
        v_x[:n], v_y[:n] = self.dst_points[:, 0], self.dst_points[:, 1]
        coef_x, coef_y = np.linalg.solve(L, v_x), np.linalg.solve(L, v_y)
        return coef_x[:n], coef_y[:n], coef_x[n:], coef_y[n:]

    def _handle(self, device_index: int) -> int:
        h = self._handles.get(device_index)
        if h is None:
            arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (self.src_points, self.weights_x, self.weights_y, self.affine_x,
                                                                        self.affine_y)]
            out = _P()
            _lib.check(_lib.lib().opd_tps_table_create(*[a.ctypes.data for a in arrs], len(self.src_points), device_index, C.byref(out)),
                       "opd_tps_table_create")
            h = self._handles[device_index] = out.value
        return h

    def __del__(self):
        try:
            for h in getattr(self, "_handles", {}).values():
                _lib.lib().opd_tps_table_destroy(h)
        except Exception:  # interpreter shutdown
            pass

    def transform_points(self, points, *, is_bbox: bool = False, with_mm: bool = False, with_bounds: bool = False):
        """[N,2] points (or [N,4] boxes) float64 CUDA tensor -> [N,2] floor pixels (optionally mm and the in-bounds flag)."""
        torch = _lib.require_cuda()
        cols = 4 if is_bbox else 2
        if points.dim() != 2 or points.shape[1] != cols or not points.is_cuda or points.dtype != torch.float64:
            raise ValueError(f"points must be a float64 CUDA tensor of shape [N,{cols}]")
        pts = points.contiguous()
        if self.distortion_corrector is not None and self.distortion_corrector.enabled:
            pts, is_bbox = self.distortion_corrector.undistort_tensor(pts, is_bbox=is_bbox), False   # :490-494
        n, dev = pts.shape[0], pts.device
        px = torch.empty((n, 2), dtype=torch.float64, device=dev)
        mm = torch.empty((n, 2), dtype=torch.float64, device=dev) if with_mm else None
        inb = torch.empty((n,), dtype=torch.uint8, device=dev) if with_bounds else None
        fm = self.floormap_config
        sx, sy = (float(fm.scale_x_mm_per_px), float(fm.scale_y_mm_per_px)) if fm else (1.0, 1.0)
        mw, mh = (float(fm.width_px), float(fm.height_px)) if fm else (float("inf"), float("inf"))
        with torch.cuda.device(dev):
            rc = _lib.lib().opd_tps_transform_f64(self._handle(dev.index), _lib.ptr(pts), int(is_bbox), n, sx, sy, mw, mh, _lib.ptr(px),
                                                  _lib.ptr(mm), _lib.ptr(inb), _lib.stream_ptr())
        _lib.check(rc, "opd_tps_transform_f64")
        out = [px] + ([mm] if with_mm else []) + ([inb] if with_bounds else [])
        return out[0] if len(out) == 1 else tuple(out)

    def _run(self, rows: np.ndarray, is_bbox: bool) -> list[PWATransformResult]:
        torch = _lib.require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(torch.device("cuda", torch.cuda.current_device()))
        px, mm, inb = (a.cpu().numpy() for a in self.transform_points(t, is_bbox=is_bbox, with_mm=True, with_bounds=True))
        has_fm = self.floormap_config is not None
        return [PWATransformResult(floor_coords_px=(float(px[i, 0]), float(px[i, 1])),
                                   floor_coords_mm=(float(mm[i, 0]), float(mm[i, 1])) if has_fm else None, is_valid=True,
                                   is_within_bounds=bool(inb[i]) if has_fm else True) for i in range(len(rows))]

    def transform_pixel(self, image_point: tuple[float, float]) -> PWATransformResult:
        return self._run(np.array([[image_point[0], image_point[1]]], dtype=np.float64), False)[0]

    def transform_detection(self, bbox: tuple[float, float, float, float]) -> PWATransformResult:
        return self._run(np.array([bbox], dtype=np.float64), True)[0]

    def transform_batch(self, bboxes: Sequence[tuple[float, float, float, float]]) -> list[PWATransformResult]:
        if len(bboxes) == 0:
            return []
        return self._run(np.array(bboxes, dtype=np.float64).reshape(-1, 4), True)

    def evaluate_training_error(self) -> dict:
        res = self._run(self.src_points, False)
        errors = [float(np.sqrt((r.floor_coords_px[0] - d[0]) ** 2 + (r.floor_coords_px[1] - d[1]) ** 2)) for r, d in zip(res, self.dst_points)]
        if not errors:
            return {"rmse": 0.0, "max_error": 0.0, "mean_error": 0.0}
        e = np.array(errors)
        return {"rmse": float(np.sqrt(np.mean(e ** 2))), "max_error": float(np.max(e)), "mean_error": float(np.mean(e))}

    def get_info(self) -> dict:
        return {"method": "thin_plate_spline", "num_points": len(self.src_points), "regularization": self.regularization,
                "training_error": self.evaluate_training_error(), "distortion_correction_enabled": self.distortion_corrector is not None,
                **({"distortion_params": self.distortion_corrector.intrinsics.distortion.to_dict()} if self.distortion_corrector is not None else {})}

    @classmethod
    def from_correspondence_file(cls, file_path: Path | str, floormap_config=None, regularization: float = 0.0,
                                 distortion_corrector=None) -> "ThinPlateSplineTransformer":
        with open(file_path, encoding="utf-8") as f:
            data = json.load(f)
        points = data.get("point_correspondences", [])
        return cls(np.array([p["src_point"] for p in points]), np.array([p["dst_point"] for p in points]), floormap_config, regularization,
                   distortion_corrector)
