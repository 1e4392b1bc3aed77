"""Per-frame zone counting (the counting step of Phase 4).

Keeps `Aggregator.get_zone_counts` / `aggregate_frame` / `export_csv` of the reference
(src/aggregation/aggregator.py:31-133).  The histogram itself is `zone_histogram_kernel`
(csrc/floor.cu); `aggregate_histogram` ingests the dense [T, Z+1] tensor that
`ZoneClassifier.count` (and the multi-GPU all-reduce) produces without building Detection objects.
Statistics / trend / peak analysis stay in the reference (out of scope).
"""

from __future__ import annotations

import csv
import logging
from collections import defaultdict

import numpy as np

from .. import _lib
from ..models.data_models import AggregationResult, Detection

logger = logging.getLogger(__name__)


class Aggregator:
    def __init__(self):
        self.results: list[AggregationResult] = []
        self._zone_data: dict[str, list[int]] = defaultdict(list)
        logger.info("Aggregator initialized")

    def get_zone_counts(self, detections: list[Detection]) -> dict[str, int]:
        """{zone_id: count}; a detection counts once in each of its zones, or once in "unclassified"
        (aggregator.py:52-75).  Key order = first appearance, like the reference's defaultdict."""
        if not detections:
            return {}
        torch = _lib.require_cuda()
        names: dict[str, int] = {}
        flat: list[int] = []
        for det in detections:
            for zid in (det.zone_ids if det.zone_ids else ("unclassified",)):
                flat.append(names.setdefault(zid, len(names)))
        dev = torch.device("cuda", torch.cuda.current_device())
        idx = torch.tensor(flat, dtype=torch.int32, device=dev)
        hist = torch.zeros((1, len(names) + 1), dtype=torch.int32, device=dev)
        _lib.check(_lib.lib().opd_zone_histogram(_lib.ptr(idx), None, None, idx.numel(), len(names), 1,
                                                 _lib.ptr(hist), _lib.stream_ptr()), "opd_zone_histogram")
        h = hist.cpu().numpy()[0]
        return {zid: int(h[i]) for zid, i in names.items()}

    def _store(self, timestamp: str, zone_counts: dict[str, int]) -> None:
        for zone_id, count in zone_counts.items():
            self.results.append(AggregationResult(timestamp=timestamp, zone_id=zone_id, count=count))
            self._zone_data[zone_id].append(count)

    def aggregate_frame(self, timestamp: str, detections: list[Detection]) -> dict[str, int]:
        """Count one frame and remember the non-zero bins (aggregator.py:31-50)."""
        zone_counts = self.get_zone_counts(detections)
        self._store(timestamp, zone_counts)
        return zone_counts

    def aggregate_histogram(self, timestamps: list[str], hist, zone_ids: list[str]) -> list[dict[str, int]]:
        """Ingest a dense [T, Z+1] histogram (rows = timestamps, last column = unclassified)."""
        h = hist.cpu().numpy() if hasattr(hist, "cpu") else np.asarray(hist)
        if h.shape != (len(timestamps), len(zone_ids) + 1):
            raise ValueError(f"hist shape {h.shape} != ({len(timestamps)}, {len(zone_ids) + 1})")
        names = [*zone_ids, "unclassified"]
        out = []
        for ts, row in zip(timestamps, h):
            counts = {names[j]: int(c) for j, c in enumerate(row) if c}
            self._store(ts, counts)
            out.append(counts)
        return out

    def export_csv(self, output_path: str, zone_ids: list[str] | None = None) -> None:
        """timestamp, <zones...>, unclassified — one row per timestamp, sorted (aggregator.py:77-133)."""
        table: dict[str, dict[str, int]] = defaultdict(dict)
        for r in self.results:
            table[r.timestamp][r.zone_id] = r.count
        if zone_ids is None:
            seen = {r.zone_id for r in self.results}
            zone_ids = sorted(z for z in seen if z.startswith("zone_"))
            if "unclassified" in seen:
                zone_ids.append("unclassified")
        elif "unclassified" not in zone_ids:
            zone_ids = [*zone_ids, "unclassified"]
        try:
            with open(output_path, "w", newline="", encoding="utf-8") as f:
                w = csv.writer(f)
                w.writerow(["timestamp", *zone_ids])
                for ts in sorted(table):
                    w.writerow([ts, *(table[ts].get(z, 0) for z in zone_ids)])
        except OSError as e:
            logger.error("Failed to export CSV: %s", e)
            raise
        logger.info("Aggregation results exported to CSV: %s (%d rows)", output_path, len(table))
