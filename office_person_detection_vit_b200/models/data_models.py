"""Records carried between Phase 2, 3 and 4.

Field names, order and defaults follow the reference's dataclasses
(src/models/data_models.py:9-38 Detection, :41-55 FrameResult, :58-70 AggregationResult) so that
objects produced here can be consumed by the reference's phases, exporters and visualisers.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class Detection:
    """One detected person.  bbox = (x, y, width, height) in camera pixels."""

    bbox: tuple[float, float, float, float]
    confidence: float
    class_id: int
    class_name: str
    camera_coords: tuple[float, float]
    floor_coords: Optional[tuple[float, float]] = None
    floor_coords_mm: Optional[tuple[float, float]] = None
    zone_ids: list[str] = field(default_factory=list)
    track_id: Optional[int] = None
    features: Optional[np.ndarray] = None
    appearance_score: Optional[float] = None
    query_index: Optional[int] = None


@dataclass
class FrameResult:
    """All detections of one sampled frame plus its zone counts."""

    frame_number: int
    timestamp: str
    detections: list[Detection]
    zone_counts: dict[str, int]


@dataclass
class AggregationResult:
    """One (timestamp, zone) count; only non-zero zones are stored (aggregator.py:44-47)."""

    timestamp: str
    zone_id: str
    count: int
