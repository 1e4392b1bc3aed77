"""Phase-3 result export straight from the device tensors of `DetectCountPipeline.run_tensors`.

Produces the reference's `coordinate_transformations.json` (src/pipeline/phases/transform.py:398-531: key names,
compact-key / precision / exclude-px options, rounding `round(float(v), precision)`, `indent=2` or None) and feeds
`Aggregator.aggregate_histogram` / `export_csv` for `zone_counts.csv` (src/aggregation/aggregator.py:77-133) WITHOUT building
one Python `Detection` per box in the hot loop: the compacted rows come to the host in four batched copies and are
formatted from NumPy arrays.  `frame_results_from_tensors` materialises `FrameResult` / `Detection` records for the
consumers that want objects (visualisation, evaluation) - the reference's schemas stay unchanged."""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Sequence

import numpy as np

from ..models import Detection, FrameResult


def _host(out: dict, transformer, zone_ids: Sequence[str]):
    """Batched device -> host copies of the compacted detection rows (+ floor coordinates from the same kernel)."""
    n_keep = out["n_keep"].cpu().numpy()
    xywh = out["det_xywh"].cpu().numpy().astype(np.float64)
    score = out["det_score"].cpu().numpy().astype(np.float64)
    foot = out["det_foot"].cpu().numpy()
    zone_idx = out["zone_idx"].cpu().numpy()
    B, Q = n_keep.shape[0], xywh.shape[1]
    px, mm = transformer.transform_points(out["det_foot"].view(-1, 2), with_mm=True)[:2]
    px = px.cpu().numpy().reshape(B, Q, 2)
    mm = mm.cpu().numpy().reshape(B, Q, 2)
    return n_keep, xywh, score, foot, px, mm, zone_idx


def _r(v: float, precision: int) -> float:
    return round(float(v), precision)   # transform.py:533-536 _round_coord


def format_coordinate_transformations(n_keep, xywh, score, foot, px, mm, zone_idx, frame_numbers: Sequence[int],
                                      timestamps: Sequence[str], transformer_info: dict, zone_ids: Sequence[str],
                                      transform_method: str = "homography", json_optimization: dict | None = None) -> dict:
    """Pure host formatting (NumPy in, dict out) of transform.py:398-531; rows r < n_keep[b] of frame b are detections."""
    opt = json_optimization or {}
    enabled = bool(opt.get("enabled", False))
    precision = opt.get("coordinate_precision", 1) if enabled else 6
    compact = bool(opt.get("compact_keys", False)) and enabled
    exclude_px = bool(opt.get("exclude_px_coords", False)) and enabled
    frames = []
    for b, (fn, ts) in enumerate(zip(frame_numbers, timestamps)):
        dets = []
        for r in range(int(n_keep[b])):
            bb, cam, fp, fm = xywh[b, r], foot[b, r], px[b, r], mm[b, r]
            if compact:
                d: dict[str, Any] = {"bb": [_r(bb[0], precision), _r(bb[1], precision), _r(bb[2], precision), _r(bb[3], precision)],
                                     "conf": _r(score[b, r], 2),
                                     "cam": [_r(cam[0], precision), _r(cam[1], precision)]}
                if not exclude_px:
                    d["floor_px"] = [_r(fp[0], precision), _r(fp[1], precision)]
                d["floor_mm"] = [_r(fm[0], precision), _r(fm[1], precision)]
                if zone_idx[b, r] >= 0:
                    d["zones"] = [zone_ids[int(zone_idx[b, r])]]
            else:
                d = {"bbox": {"x": _r(bb[0], precision), "y": _r(bb[1], precision), "width": _r(bb[2], precision),
                              "height": _r(bb[3], precision)},
                     "confidence": _r(score[b, r], 3),
                     "camera_coords": {"x": _r(cam[0], precision), "y": _r(cam[1], precision)}}
                if not exclude_px:
                    d["floor_coords_px"] = {"x": _r(fp[0], precision), "y": _r(fp[1], precision)}
                d["floor_coords_mm"] = {"x": _r(fm[0], precision), "y": _r(fm[1], precision)}
                if zone_idx[b, r] >= 0:
                    d["zone_ids"] = [zone_ids[int(zone_idx[b, r])]]
            dets.append(d)
        frames.append({("idx" if compact else "frame_number"): fn, ("ts" if compact else "timestamp"): ts,
                       ("det" if compact else "detections"): dets})
    info = dict(transformer_info or {})
    if compact and info:
        info = {"method": info.get("method", transform_method), "points": info.get("num_points", 0),
                "triangles": info.get("num_triangles", 0)}
    return {("method" if compact else "transform_method"): transform_method, ("info" if compact else "transformer_info"): info,
            "frames": frames}


def dumps_coordinate_transformations(data: dict, json_optimization: dict | None = None) -> str:
    """The exact text the reference writes (indent 2, or none with compact keys; ensure_ascii=False)."""
    opt = json_optimization or {}
    compact = bool(opt.get("compact_keys", False)) and bool(opt.get("enabled", False))
    return json.dumps(data, indent=None if compact else 2, ensure_ascii=False, default=str)


def coordinate_transformations_dict(out: dict, frame_numbers: Sequence[int], timestamps: Sequence[str], transformer,
                                    zone_ids: Sequence[str], transform_method: str = "homography",
                                    json_optimization: dict | None = None) -> dict:
    """The dict the reference dumps to coordinate_transformations.json, built from the pipeline's tensors."""
    n_keep, xywh, score, foot, px, mm, zone_idx = _host(out, transformer, zone_ids)
    info = transformer.get_info() if transformer is not None else {}
    return format_coordinate_transformations(n_keep, xywh, score, foot, px, mm, zone_idx, frame_numbers, timestamps, info,
                                             zone_ids, transform_method, json_optimization)


def export_results(out: dict, frame_numbers: Sequence[int], timestamps: Sequence[str], transformer, zone_ids: Sequence[str],
                   output_path: str | Path, transform_method: str = "homography", json_optimization: dict | None = None) -> Path:
    """Write <output_path>/coordinate_transformations.json exactly as TransformPhase.export_results does."""
    data = coordinate_transformations_dict(out, frame_numbers, timestamps, transformer, zone_ids, transform_method,
                                           json_optimization)
    path = Path(output_path) / "coordinate_transformations.json"
    path.write_text(dumps_coordinate_transformations(data, json_optimization), encoding="utf-8")
    return path


def frame_results_from_tensors(out: dict, frame_numbers: Sequence[int], timestamps: Sequence[str], transformer,
                               zone_ids: Sequence[str]) -> list[FrameResult]:
    """`FrameResult` / `Detection` records (src/models/data_models.py) for consumers that want objects; zone_counts
    are filled from the device histogram (rows of `out["hist"]` belonging to these frames must be passed in order)."""
    n_keep, xywh, score, foot, px, mm, zone_idx = _host(out, transformer, zone_ids)
    names = [*zone_ids, "unclassified"]
    hist = out["hist"].cpu().numpy()
    res = []
    for b, (fn, ts) in enumerate(zip(frame_numbers, timestamps)):
        dets = []
        for r in range(int(n_keep[b])):
            z = int(zone_idx[b, r])
            dets.append(Detection(bbox=tuple(float(v) for v in xywh[b, r]), confidence=float(score[b, r]), class_id=1,
                                  class_name="person", camera_coords=(float(foot[b, r, 0]), float(foot[b, r, 1])),
                                  floor_coords=(float(px[b, r, 0]), float(px[b, r, 1])),
                                  floor_coords_mm=(float(mm[b, r, 0]), float(mm[b, r, 1])),
                                  zone_ids=[zone_ids[z]] if z >= 0 else []))
        row = hist[hist.shape[0] - len(frame_numbers) + b] if hist.shape[0] >= len(frame_numbers) else hist[b]
        res.append(FrameResult(frame_number=fn, timestamp=ts, detections=dets,
                               zone_counts={names[j]: int(c) for j, c in enumerate(row) if c}))
    return res
