"""Phase 2 / 3 / 4 `execute` functions of the reference's pipeline, batched onto the GPU engines.

SURVEY.md §8a rows a1 / a18 / a20: the reference runs three per-frame Python loops

    DetectionPhase.execute      src/pipeline/phases/detection.py:56-133   one detector call per frame, per-frame try/except -> []
    TransformPhase.execute      src/pipeline/phases/transform.py:257-330  transform_batch + classify per detection, FrameResult records
    AggregationPhase.execute    src/pipeline/phases/aggregation.py:26-91  aggregate_frame per frame, zone_counts write-back, CSV

These classes keep the reference's phase surface - `__init__(config, logger)`, `initialize()`, `execute(...)` with the same
arguments and return types, `export_results`, `cleanup()`; `config` is anything with `.get(dotted_key, default)` (the reference's
ConfigManager) or a plain nested dict - and the same records (`Detection`, `FrameResult`), so the orchestrator, exporters and
visualisers around them do not change.  What changes is the shape of the work: Phase 2 is ONE `detect_batch` over all sampled
frames (device batches of `detection.batch_size`, failures isolated per frame like the reference's try/except), Phase 3 is ONE
projection launch and ONE classification launch over every detection of every frame, Phase 4 counts from the records.

Outputs are checked against files the reference's own, unmodified phases wrote for the same detections
(tests/golden/make_phase_golden.py -> tests/golden/phase_golden/, tests/test_phases_gpu.py).

Not carried over (out of scope): detection-image saving, the statistics / trend / peak report of Phase 4 (logging
only in the reference), OutputPolicy.  `output_policy` is accepted and ignored."""

from __future__ import annotations

import logging
from pathlib import Path
from typing import Any, Sequence

import numpy as np

from .aggregation import Aggregator
from .detection import ViTDetector
from .export.results import dumps_coordinate_transformations
from .models import Detection, FrameResult
from .transform import FloorMapConfig, HomographyTransformer
from .zone import ZoneClassifier


def _cfg(config: Any, key: str, default: Any = None) -> Any:
    """`config.get("a.b", default)` on the reference's ConfigManager, or the same lookup on a nested dict."""
    if config is None:
        return default
    if not isinstance(config, dict):
        return config.get(key, default)
    node: Any = config
    for part in key.split("."):
        if not isinstance(node, dict) or part not in node:
            return default
        node = node[part]
    return node


class _Phase:
    def __init__(self, config: Any, logger: logging.Logger | None = None):
        self.config = config
        self.logger = logger or logging.getLogger(__name__)

    def log_phase_start(self, phase_name: str) -> None:
        self.logger.info("=" * 80)
        self.logger.info(phase_name)
        self.logger.info("=" * 80)

    def cleanup(self) -> None:
        pass


class DetectionPhase(_Phase):
    """Phase 2 (detection.py:19-133) on ViTDetector: all sampled frames in one batched call."""

    def __init__(self, config: Any, logger: logging.Logger | None = None, detector: ViTDetector | None = None):
        super().__init__(config, logger)
        self.detector = detector
        self.output_path: Path | None = None
        self.with_features = bool(_cfg(config, "detection.with_features", True))   # the reference calls detect_with_features

    def initialize(self) -> None:
        self.log_phase_start("フェーズ2: DETR人物検出 (B200)")
        if self.detector is None:
            self.detector = ViTDetector(
                model_name=_cfg(self.config, "detection.model_name", "facebook/detr-resnet-50"),
                confidence_threshold=_cfg(self.config, "detection.confidence_threshold", 0.5),
                device=_cfg(self.config, "detection.device"),
                batch_size=int(_cfg(self.config, "detection.batch_size", 64)))
        if self.detector.model is None:
            self.detector.load_model()

    def execute(self, sample_frames: Sequence[tuple[int, str, np.ndarray]], output_policy: Any = None
                ) -> list[tuple[int, str, list[Detection]]]:
        """[(frame_num, timestamp, frame)] -> [(frame_num, timestamp, detections)] (detection.py:56-133).  A frame whose detection
        fails yields [] and an error log; the others are unaffected (detection.py:124-127)."""
        if self.detector is None:
            raise RuntimeError("検出器が初期化されていません。initialize()を先に呼び出してください。")
        frames = [f for _, _, f in sample_frames]
        try:
            if self.with_features:
                dets, feats = self.detector.detect_batch_with_features(frames)
            else:
                dets, feats = self.detector.detect_batch(frames), None
        except Exception as e:   # nothing a single frame can cause (those are isolated inside detect_batch): log, return empties
            self.logger.error(f"検出処理に失敗しました: {e}", exc_info=True)
            dets, feats = [[] for _ in frames], None
        results = []
        for i, (frame_num, timestamp, _) in enumerate(sample_frames):
            if feats is not None and feats[i].shape[0] != len(dets[i]):
                self.logger.warning(f"フレーム #{frame_num}: 特徴量数と検出数が不一致 (features={feats[i].shape[0]}, detections={len(dets[i])})")
            results.append((frame_num, timestamp, dets[i]))
            self.logger.info(f"フレーム #{frame_num} ({timestamp}): {len(dets[i])}人検出")
        return results

    def cleanup(self) -> None:
        self.detector = None


class TransformPhase(_Phase):
    """Phase 3 (transform.py:35-545) for `transform.method: homography`: every detection of every frame in one launch each for
    the projection and the zone test."""

    SUPPORTED_METHODS = ("homography",)

    def __init__(self, config: Any, logger: logging.Logger | None = None):
        super().__init__(config, logger)
        self.transformer: HomographyTransformer | None = None
        self.zone_classifier: ZoneClassifier | None = None
        self.transform_method = "homography"

    def _create_floormap_config(self) -> FloorMapConfig:
        fm = _cfg(self.config, "floormap", {}) or {}
        return FloorMapConfig(
            width_px=int(fm.get("image_width", 1878)), height_px=int(fm.get("image_height", 1369)),
            origin_x_px=float(fm.get("image_origin_x", 7.0)), origin_y_px=float(fm.get("image_origin_y", 9.0)),
            scale_x_mm_per_px=float(fm.get("image_x_mm_per_pixel", 28.1926406926406)),
            scale_y_mm_per_px=float(fm.get("image_y_mm_per_pixel", 28.241430700447)))

    def initialize(self) -> None:
        method = (_cfg(self.config, "transform", {}) or {}).get("method", "homography")
        if method not in self.SUPPORTED_METHODS:
            # the piecewise-affine / TPS transformers exist (transform/piecewise_affine.py) but this phase wires the homography path
            self.logger.warning(f"transform method '{method}' is not wired into this phase, falling back to 'homography'")
        self.log_phase_start(f"フェーズ3: 座標変換とゾーン判定 ({self.transform_method})")
        matrix = (_cfg(self.config, "homography", {}) or {}).get("matrix")
        if matrix is None:
            raise ValueError("homography.matrix が設定されていません")          # transform.py:144-145
        H = np.array(matrix, dtype=np.float64)
        if H.shape != (3, 3):
            raise ValueError(f"ホモグラフィ行列は3x3である必要があります: {H.shape}")
        self.transformer = HomographyTransformer(H, self._create_floormap_config())
        zones = _cfg(self.config, "zones", []) or []
        if not zones:
            self.logger.warning("ゾーン定義が設定されていません")
        self.zone_classifier = ZoneClassifier(zones, allow_overlap=False)     # transform.py:254
        self.logger.info(f"ZoneClassifier initialized with {len(zones)} zones.")

    def execute(self, detection_results: Sequence[tuple[int, str, list[Detection]]]) -> list[FrameResult]:
        """[(frame_num, timestamp, detections)] -> [FrameResult] (transform.py:257-330): floor_coords, floor_coords_mm, camera_coords
        and zone_ids are written into the Detection records, zone_counts stays {} for Phase 4."""
        if self.transformer is None or self.zone_classifier is None:
            raise RuntimeError("Not initialized. Call initialize() first.")
        flat = [d for _, _, dets in detection_results for d in dets]
        out_of_bounds = classified = 0
        if flat:
            results = self.transformer.transform_batch([d.bbox for d in flat])                 # one launch
            zones = self.zone_classifier.classify_batch([r.floor_coords_px for r in results])   # one launch
            for d, r, z in zip(flat, results, zones):
                d.floor_coords = r.floor_coords_px
                d.floor_coords_mm = r.floor_coords_mm
                if d.bbox:
                    x, y, w, h = d.bbox
                    d.camera_coords = (x + w / 2.0, y + h)                                      # transform.py:347-349
                d.zone_ids = z
                out_of_bounds += not r.is_within_bounds
                classified += bool(z)
        frame_results = [FrameResult(frame_number=fn, timestamp=ts, detections=dets, zone_counts={})
                         for fn, ts, dets in detection_results]
        total = len(flat)
        if total:
            self.logger.info("=" * 80)
            self.logger.info(f"Phase 3 Statistics ({self.transform_method}):")
            self.logger.info(f"  Total Detections: {total}")
            self.logger.info(f"  Transform Success: {total} (100.0%)")
            self.logger.info("  Transform Errors: 0 (0.0%)")
            self.logger.info(f"  Out of Bounds: {out_of_bounds} ({out_of_bounds / total * 100:.1f}%)")
            self.logger.info(f"  Zone Classified: {classified} ({classified / total * 100:.1f}%)")
            self.logger.info("=" * 80)
        return frame_results

    def export_results(self, frame_results: Sequence[FrameResult], output_path: Path) -> None:
        """coordinate_transformations.json as transform.py:398-531 writes it (key names, precision and compact-key options)."""
        opt = _cfg(self.config, "output.json_optimization", {}) or {}
        enabled = bool(opt.get("enabled", False))
        precision = opt.get("coordinate_precision", 1) if enabled else 6
        compact = bool(opt.get("compact_keys", False)) and enabled
        exclude_px = bool(opt.get("exclude_px_coords", False)) and enabled

        def pair(v, keys=("x", "y")):
            vals = [round(c, precision) for c in v]
            return vals if compact else dict(zip(keys, vals))

        frames = []
        for fr in frame_results:
            dets = []
            for d in fr.detections:
                e: dict[str, Any] = {("bb" if compact else "bbox"): pair(d.bbox, ("x", "y", "width", "height")),
                                     ("conf" if compact else "confidence"): round(d.confidence, 2 if compact else 3)}
                if d.camera_coords is not None:
                    e["cam" if compact else "camera_coords"] = pair(d.camera_coords)
                if d.floor_coords is not None and not exclude_px:
                    e["floor_px" if compact else "floor_coords_px"] = pair(d.floor_coords)
                if d.floor_coords_mm is not None:
                    e["floor_mm" if compact else "floor_coords_mm"] = pair(d.floor_coords_mm)
                if d.zone_ids:
                    e["zones" if compact else "zone_ids"] = d.zone_ids
                if getattr(d, "track_id", None) is not None:
                    e["id" if compact else "track_id"] = d.track_id
                dets.append(e)
            frames.append({("idx" if compact else "frame_number"): fr.frame_number, ("ts" if compact else "timestamp"): fr.timestamp,
                           ("det" if compact else "detections"): dets})
        info = self.transformer.get_info() if self.transformer else {}
        if compact and info:
            info = {"method": info.get("method", self.transform_method), "points": info.get("num_points", 0),
                    "triangles": info.get("num_triangles", 0)}
        data = {("method" if compact else "transform_method"): self.transform_method, ("info" if compact else "transformer_info"): info,
                "frames": frames}
        path = Path(output_path) / "coordinate_transformations.json"
        try:
            path.write_text(dumps_coordinate_transformations(data, opt), encoding="utf-8")
            self.logger.info(f"Saved coordinate transformations to {path}")
        except OSError as e:
            self.logger.error(f"Failed to save JSON: {e}")

    def cleanup(self) -> None:
        self.transformer = None
        self.zone_classifier = None


class AggregationPhase(_Phase):
    """Phase 4's counting step (aggregation.py:26-91): per-frame zone counts written back into the FrameResults, zone_counts.csv
    in the configured zone order.  (The statistics / trend / peak log lines of the reference are reporting, not on this path.)"""

    def execute(self, frame_results: Sequence[FrameResult], output_path: Path) -> Aggregator:
        self.log_phase_start("フェーズ4: 集計とレポート生成")
        aggregator = Aggregator()
        for fr in frame_results:
            fr.zone_counts = aggregator.aggregate_frame(fr.timestamp, fr.detections)
        csv_path = Path(output_path) / "zone_counts.csv"
        zones = _cfg(self.config, "zones", []) or []
        zone_ids = [z["id"] for z in zones] if zones else None
        aggregator.export_csv(str(csv_path), zone_ids=zone_ids)
        self.logger.info(f"集計結果をCSVに出力しました: {csv_path}")
        return aggregator
