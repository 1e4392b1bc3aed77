"""Floormap geometry (pixel size, origin offset, mm-per-pixel scales).

Same fields, defaults and `from_config` keys as the reference's FloorMapConfig
(src/transform/floormap_config.py:13-70).
"""

from __future__ import annotations

from dataclasses import dataclass


@dataclass
class FloorMapConfig:
    width_px: int = 1878
    height_px: int = 1369
    origin_x_px: float = 7.0
    origin_y_px: float = 9.0
    scale_x_mm_per_px: float = 28.1926406926406
    scale_y_mm_per_px: float = 28.241430700447

    @classmethod
    def from_config(cls, config: dict) -> "FloorMapConfig":
        """Build from the `floormap` section of config.yaml (config.yaml:208-223)."""
        g = config.get
        return cls(
            width_px=int(g("image_width", 1878)),
            height_px=int(g("image_height", 1369)),
            origin_x_px=float(g("image_origin_x", 7.0)),
            origin_y_px=float(g("image_origin_y", 9.0)),
            scale_x_mm_per_px=float(g("image_x_mm_per_pixel", 28.1926406926406)),
            scale_y_mm_per_px=float(g("image_y_mm_per_pixel", 28.241430700447)),
        )

    @property
    def scale_x_m_per_px(self) -> float:
        return self.scale_x_mm_per_px / 1000.0

    @property
    def scale_y_m_per_px(self) -> float:
        return self.scale_y_mm_per_px / 1000.0

    @property
    def scale_x_px_per_m(self) -> float:
        return 1000.0 / self.scale_x_mm_per_px

    @property
    def scale_y_px_per_m(self) -> float:
        return 1000.0 / self.scale_y_mm_per_px
