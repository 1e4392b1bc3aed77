"""Camera pixel -> floormap projection on the GPU.

Keeps the surface of the reference's HomographyTransformer (src/transform/homography.py:40-213):
`transform_pixel`, `transform_detection`, `transform_batch`, `get_info`, the `TransformResult` record and
the constructor's ValueError behaviour (:66-91).  The arithmetic (foot point :166-169, H·p and perspective
divide :172-175, bounds :181, mm scale :183-186) runs in `floor_exact_kernel` (csrc/floor.cu) in float64.

`transform_points` is the tensor entry the north star names: [N,2] float32/float64 CUDA tensor in,
[N,2] floor pixels out, no Python objects.
"""

from __future__ import annotations

import logging
from dataclasses import dataclass
from typing import TYPE_CHECKING, Sequence

import numpy as np

from .. import _lib

if TYPE_CHECKING:
    from .floormap_config import FloorMapConfig

logger = logging.getLogger(__name__)


@dataclass
class TransformResult:
    """Result of projecting one point (homography.py:21-37)."""

    floor_coords_px: tuple[float, float] | None = None
    floor_coords_mm: tuple[float, float] | None = None
    is_valid: bool = False
    error_reason: str | None = None
    is_within_bounds: bool = False


class HomographyTransformer:
    """3x3 homography from camera pixels to floormap pixels, evaluated on the GPU."""

    def __init__(self, homography_matrix, floormap_config: "FloorMapConfig"):
        self.H = self._validate_matrix(homography_matrix)
        self.floormap_config = floormap_config
        self._empty_zones = None  # lazily created Z=0 zone table (projection-only launches)
        logger.debug("HomographyTransformer initialized with matrix:\n%s", self.H)

    # -- construction ---------------------------------------------------------------------------------
    @staticmethod
    def _validate_matrix(matrix) -> np.ndarray:
        """Shape / singularity checks with the reference's messages (homography.py:66-91)."""
        H = np.array(matrix, dtype=np.float64)
        if H.shape != (3, 3):
            raise ValueError(f"ホモグラフィ行列は3x3である必要があります: {H.shape}")
        det = np.linalg.det(H)
        if abs(det) < 1e-10:
            raise ValueError(f"ホモグラフィ行列が特異行列です（行列式={det}）")
        cond = np.linalg.cond(H)
        if cond > 1e12:
            logger.warning("ホモグラフィ行列の条件数が大きい: %s", cond)
        return H

    def floor_params(self, input_is_bbox: bool = False, skip_projection: bool = False) -> "_lib.FloorParams":
        fm = self.floormap_config
        p = _lib.FloorParams()
        for i, v in enumerate(self.H.reshape(-1)):
            p.H[i] = float(v)
        p.scale_x_mm = float(fm.scale_x_mm_per_px)
        p.scale_y_mm = float(fm.scale_y_mm_per_px)
        p.map_w_px = float(fm.width_px)
        p.map_h_px = float(fm.height_px)
        p.input_is_bbox = int(input_is_bbox)
        p.skip_projection = int(skip_projection)
        return p

    def _zones0(self):
        if self._empty_zones is None:
            from ..zone.zone_classifier import ZoneTable

            self._empty_zones = ZoneTable([], allow_overlap=False)
        return self._empty_zones

    # -- tensor entries -------------------------------------------------------------------------------
    def transform_points(self, points, *, is_bbox: bool = False, with_mm: bool = False, with_bounds: bool = False):
        """[N,2] points (or [N,4] boxes with is_bbox) CUDA tensor -> [N,2] floor pixels (same dtype).

        Optionally also returns floor mm and the in-bounds flag: (px[, mm][, within])."""
        torch = _lib.require_cuda()
        cols = 4 if is_bbox else 2
        if points.dim() != 2 or points.shape[1] != cols or not points.is_cuda:
            raise ValueError(f"points must be a CUDA tensor of shape [N,{cols}]")
        if points.dtype not in (torch.float32, torch.float64):
            raise ValueError("points must be float32 or float64")
        pts = points.contiguous()
        n = pts.shape[0]
        px = torch.empty((n, 2), dtype=pts.dtype, device=pts.device)
        mm = torch.empty((n, 2), dtype=pts.dtype, device=pts.device) if with_mm else None
        inb = torch.empty((n,), dtype=torch.uint8, device=pts.device) if with_bounds else None
        fn = (_lib.lib().opd_floor_project_classify_count_f64 if pts.dtype == torch.float64
              else _lib.lib().opd_floor_project_classify_count_f32)
        params = self.floor_params(input_is_bbox=is_bbox)
        with torch.cuda.device(pts.device):
            zt = self._zones0().handle(pts.device.index)
            _lib.check(fn(params, zt, _lib.ptr(pts), None, n, 1, _lib.ptr(px), _lib.ptr(mm), _lib.ptr(inb),
                          None, None, None, _lib.stream_ptr()), "transform_points")
        out = [px]
        if with_mm:
            out.append(mm)
        if with_bounds:
            out.append(inb)
        return out[0] if len(out) == 1 else tuple(out)

    # -- reference surface ----------------------------------------------------------------------------
    def _get_foot_point(self, bbox: tuple[float, float, float, float]) -> tuple[float, float]:
        x, y, w, h = bbox
        return (x + w / 2, y + h)

    def _run(self, rows: np.ndarray, is_bbox: bool) -> list[TransformResult]:
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float64)).to(dev)
        px, mm, inb = self.transform_points(t, is_bbox=is_bbox, with_mm=True, with_bounds=True)
        px, mm, inb = px.cpu().numpy(), mm.cpu().numpy(), inb.cpu().numpy()
        return [
            TransformResult(
                is_valid=True,  # the reference never marks a projected point invalid (homography.py:188-195)
                floor_coords_px=(float(px[i, 0]), float(px[i, 1])),
                floor_coords_mm=(float(mm[i, 0]), float(mm[i, 1])),
                is_within_bounds=bool(inb[i]),
            )
            for i in range(len(rows))
        ]

    def transform_pixel(self, image_point: tuple[float, float]) -> TransformResult:
        """Project one camera point (homography.py:105-133)."""
        return self._run(np.array([[image_point[0], image_point[1]]], dtype=np.float64), False)[0]

    def transform_detection(self, bbox: tuple[float, float, float, float]) -> TransformResult:
        """Project the foot point of one (x, y, w, h) box (homography.py:135-148)."""
        return self._run(np.array([bbox], dtype=np.float64), True)[0]

    def transform_batch(self, bboxes: Sequence[tuple[float, float, float, float]]) -> list[TransformResult]:
        """Project the foot points of many boxes in one launch (homography.py:150-197)."""
        if len(bboxes) == 0:
            return []
        return self._run(np.array(bboxes, dtype=np.float64).reshape(-1, 4), True)

    def get_info(self) -> dict:
        fm = self.floormap_config
        return {
            "method": "homography",
            "matrix": self.H.tolist(),
            "floormap_size": (fm.width_px, fm.height_px),
            "scale_mm_per_px": (fm.scale_x_mm_per_px, fm.scale_y_mm_per_px),
        }
