"""Polygon zone classification on the GPU.

Keeps the surface of the reference's ZoneClassifier (src/zone/zone_classifier.py:8-243): constructor
validation and its ValueErrors (:44-112), `classify` (:114-149), `classify_batch` (:151-160),
`classify_with_unclassified`, `get_zone_info`, `get_all_zone_ids`, `get_zone_count`.  The ray casting
(:162-197) and the priority pick (:138-146) run in csrc/floor.cu.

Tensor entries (north star): `classify_points(points)` -> int32 zone index per point (-1 = no zone) and
`count(points, slot, T)` -> [T, Z+1] int32 histogram (last column = "unclassified").
"""

from __future__ import annotations

import ctypes as C
import logging
import math
from typing import Sequence

import numpy as np

from .. import _lib

logger = logging.getLogger(__name__)


class ZoneTable:
    """Device-side polygon table + grid accelerator (one handle per CUDA device)."""

    def __init__(self, zones: list[dict], allow_overlap: bool):
        if len(zones) > _lib.OPD_MAX_ZONES:
            raise ValueError(f"one device table holds at most {_lib.OPD_MAX_ZONES} zones, got {len(zones)} (ZoneClassifier splits "
                             "larger zone lists into groups)")
        self.allow_overlap = bool(allow_overlap)
        self.Z = len(zones)
        offs = [0]
        verts: list[tuple[float, float]] = []
        for z in zones:
            verts.extend(z["polygon"])
            offs.append(len(verts))
        self._verts = np.ascontiguousarray(np.array(verts, dtype=np.float64).reshape(-1, 2))
        self._offs = np.ascontiguousarray(np.array(offs, dtype=np.int32))
        self._prio = np.ascontiguousarray(
            np.array([math.inf if z.get("priority") is None else float(z["priority"]) for z in zones],
                     dtype=np.float64))
        self._handles: dict[int, int] = {}

    def handle(self, device_index: int | None = None) -> int:
        torch = _lib.require_cuda()
        dev = torch.cuda.current_device() if device_index is None else device_index
        h = self._handles.get(dev)
        if h is None:
            out = C.c_void_p()
            rc = _lib.lib().opd_zone_table_create(
                self._verts.ctypes.data if self.Z else None, self._offs.ctypes.data if self.Z else None,
                self._prio.ctypes.data if self.Z else None, self.Z, int(self.allow_overlap), dev, C.byref(out))
            _lib.check(rc, "opd_zone_table_create")
            h = out.value
            self._handles[dev] = h
        return h

    def info(self, device_index: int | None = None) -> dict:
        gw, gh, nb, nc = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        _lib.check(_lib.lib().opd_zone_table_info(self.handle(device_index), C.byref(gw), C.byref(gh), C.byref(nb),
                                                  C.byref(nc)))
        return {"grid_w": gw.value, "grid_h": gh.value, "boundary_cells": nb.value, "classes": nc.value}

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().opd_zone_table_destroy(h)
        except Exception:  # interpreter shutdown
            pass


class ZoneClassifier:
    """Which zone(s) does a floormap point fall in?  Ray casting, evaluated on the GPU."""

    def __init__(self, zones: list[dict], allow_overlap: bool = True):
        self.allow_overlap = allow_overlap
        self.zones = self._validate_zones(zones)
        # The reference sets no limit on the number of zones (zone_classifier.py:44-112).  One device table holds 64 (its masks are
        # 64-bit words): longer lists are split into groups of 64 in declaration order, every group answers with its own table and
        # opd_zone_combine_groups keeps the best group winner by the global (priority or +inf, declaration order) rank.
        M = _lib.OPD_MAX_ZONES
        self._tables = [ZoneTable(self.zones[g:g + M], allow_overlap) for g in range(0, max(len(self.zones), 1), M)]
        self._table = self._tables[0]
        order = sorted(range(len(self.zones)),
                       key=lambda i: (math.inf if self.zones[i]["priority"] is None else self.zones[i]["priority"], i))
        self._rank = np.empty(len(self.zones), dtype=np.int32)
        self._rank[order] = np.arange(len(self.zones), dtype=np.int32)
        self._rank_dev: dict = {}
        self._ids = [z["id"] for z in self.zones]
        logger.info("ZoneClassifierを初期化しました。ゾーン数: %d, allow_overlap=%s", len(self.zones), self.allow_overlap)

    # -- validation: same rules and messages as zone_classifier.py:44-112 -------------------------------
    @staticmethod
    def _validate_zones(zones) -> list[dict]:
        if not isinstance(zones, list):
            raise ValueError("zonesはリストである必要があります。")
        seen: set = set()
        out: list[dict] = []
        for i, zone in enumerate(zones):
            if not isinstance(zone, dict):
                raise ValueError(f"zones[{i}]は辞書である必要があります。")
            if "id" not in zone:
                raise ValueError(f"zones[{i}]には'id'が必要です。")
            zid = zone["id"]
            if zid in seen:
                raise ValueError(f"重複したゾーンID: {zid}")
            seen.add(zid)
            if "polygon" not in zone:
                raise ValueError(f"zones[{i}]には'polygon'が必要です。")
            polygon = zone["polygon"]
            if not isinstance(polygon, list) or len(polygon) < 3:
                raise ValueError(f"zones[{i}].polygonは少なくとも3つの頂点が必要です。")
            pts: list[tuple[float, float]] = []
            for j, point in enumerate(polygon):
                if not isinstance(point, (list, tuple)) or len(point) != 2:
                    raise ValueError(f"zones[{i}].polygon[{j}]は[x, y]形式である必要があります。")
                try:
                    pts.append((float(point[0]), float(point[1])))
                except (ValueError, TypeError) as e:
                    raise ValueError(f"zones[{i}].polygon[{j}]の座標は数値である必要があります。") from e
            entry = {"id": zid, "name": zone.get("name", zid), "polygon": pts, "priority": None, "_order": i}
            if zone.get("priority") is not None:
                try:
                    entry["priority"] = float(zone["priority"])
                except (TypeError, ValueError) as e:
                    raise ValueError(f"zones[{i}].priority は数値である必要があります。") from e
            out.append(entry)
        return out

    # -- tensor entries ---------------------------------------------------------------------------------
    @property
    def table(self) -> ZoneTable:
        """The device table (the first group's when there are more than 64 zones)."""
        return self._table

    @property
    def grouped(self) -> bool:
        return len(self._tables) > 1

    def _rank_on(self, device):
        torch = _lib.require_cuda()
        r = self._rank_dev.get(device)
        if r is None:
            r = self._rank_dev[device] = torch.from_numpy(self._rank).to(device)
        return r

    def _launch_grouped(self, pts, *, want_idx=False, want_mask=False, hist=None, slot=None, T=1, transformer=None, idx_out=None):
        """More than 64 zones: one launch per group of 64, then the combine kernel (and the histogram kernel for counts)."""
        torch = _lib.require_cuda()
        n, G, Z = pts.shape[0], len(self._tables), len(self.zones)
        idx_g = torch.empty((G, n), dtype=torch.int32, device=pts.device)
        masks = torch.empty((n, G), dtype=torch.int64, device=pts.device) if (want_mask or (hist is not None and self.allow_overlap)) else None
        for g, table in enumerate(self._tables):
            _, m = self._launch(pts, table=table, want_mask=masks is not None, transformer=transformer, idx_out=idx_g[g])
            if masks is not None:
                masks[:, g] = m
        idx = idx_out if idx_out is not None else torch.empty((n,), dtype=torch.int32, device=pts.device)
        if slot is not None:
            slot = slot.to(device=pts.device, dtype=torch.int32).contiguous()
        with torch.cuda.device(pts.device):
            _lib.check(_lib.lib().opd_zone_combine_groups(_lib.ptr(idx_g), _lib.ptr(self._rank_on(pts.device)), G, n, _lib.ptr(idx), None,
                                                          _lib.stream_ptr()), "opd_zone_combine_groups")
            if slot is not None:     # rows outside [0, T) are unused rows: not classified, not counted
                idx.masked_fill_((slot < 0) | (slot >= T), -1)
            if hist is not None:
                if not self.allow_overlap:
                    _lib.check(_lib.lib().opd_zone_histogram(_lib.ptr(idx), None, _lib.ptr(slot), n, Z, T, _lib.ptr(hist), _lib.stream_ptr()),
                               "opd_zone_histogram")
                else:
                    # one count per containing zone (aggregator.py:66-69): per-group mask histograms into the group's columns, the
                    # "unclassified" column from the combined index
                    M = _lib.OPD_MAX_ZONES
                    tmp = torch.zeros((T, Z + 1), dtype=torch.int32, device=pts.device)
                    _lib.check(_lib.lib().opd_zone_histogram(_lib.ptr(idx), None, _lib.ptr(slot), n, Z, T, _lib.ptr(tmp), _lib.stream_ptr()),
                               "opd_zone_histogram")
                    hist[:, Z] += tmp[:, Z]
                    for g, table in enumerate(self._tables):
                        hg = torch.zeros((T, table.Z + 1), dtype=torch.int32, device=pts.device)
                        mg = masks[:, g].contiguous()
                        _lib.check(_lib.lib().opd_zone_histogram(None, _lib.ptr(mg), _lib.ptr(slot), n, table.Z, T, _lib.ptr(hg),
                                                                 _lib.stream_ptr()), "opd_zone_histogram")
                        hist[:, g * M:g * M + table.Z] += hg[:, :table.Z]
        return (idx if (want_idx or idx_out is not None) else None), masks

    def _launch(self, pts, *, want_idx=False, want_mask=False, hist=None, slot=None, T=1, transformer=None, idx_out=None, table=None):
        torch = _lib.require_cuda()
        if table is None and self.grouped:
            return self._launch_grouped(pts.contiguous(), want_idx=want_idx, want_mask=want_mask, hist=hist, slot=slot, T=T,
                                        transformer=transformer, idx_out=idx_out)
        table = table or self._table
        if pts.dim() != 2 or pts.shape[1] != 2 or not pts.is_cuda:
            raise ValueError("points must be a CUDA tensor of shape [N,2]")
        if pts.dtype not in (torch.float32, torch.float64):
            raise ValueError("points must be float32 or float64")
        pts = pts.contiguous()
        n = pts.shape[0]
        if idx_out is not None:
            if tuple(idx_out.shape) != (n,) or idx_out.dtype != torch.int32 or not idx_out.is_contiguous() or idx_out.device != pts.device:
                raise ValueError("index_out must be a contiguous int32 [N] tensor on the points' device")
            idx = idx_out
        else:
            idx = torch.empty((n,), dtype=torch.int32, device=pts.device) if want_idx else None
        mask = torch.empty((n,), dtype=torch.int64, device=pts.device) if want_mask else None
        if transformer is None:
            params = _lib.FloorParams()
            for i, v in enumerate((1, 0, 0, 0, 1, 0, 0, 0, 1)):
                params.H[i] = float(v)
            params.skip_projection = 1
        else:
            params = transformer.floor_params()
        if slot is not None:
            slot = slot.to(device=pts.device, dtype=torch.int32).contiguous()
        fn = (_lib.lib().opd_floor_project_classify_count_f64 if pts.dtype == torch.float64
              else _lib.lib().opd_floor_project_classify_count_f32)
        with torch.cuda.device(pts.device):
            zt = table.handle(pts.device.index)
            _lib.check(fn(params, zt, _lib.ptr(pts), _lib.ptr(slot), n, T, None, None, None, _lib.ptr(idx),
                          _lib.ptr(mask), _lib.ptr(hist), _lib.stream_ptr()), "classify")
        return idx, mask

    def classify_points(self, points, transformer=None):
        """[N,2] CUDA tensor -> int32 [N] zone index in declaration order, -1 = no zone.

        With `transformer` the points are camera pixels and are projected first (fused kernel).
        With allow_overlap=True the index is the highest-priority containing zone; use `classify_masks`
        for the full set."""
        return self._launch(points, want_idx=True, transformer=transformer)[0]

    def classify_masks(self, points, transformer=None):
        """[N,2] CUDA tensor -> int64 [N] bit mask of containing zones (bit z = zone z); with more than 64 zones int64 [N, G], word g
        holding zones 64 g .. 64 g + 63."""
        return self._launch(points, want_mask=True, transformer=transformer)[1]

    def count(self, points, slot=None, num_slots: int = 1, transformer=None, out=None, return_index: bool = False,
              index_out=None):
        """Per-timestamp zone histogram: int32 [num_slots, Z+1]; column Z counts unclassified points
        (aggregator.py:64-75).  `slot[i]` is the timestamp row of point i (None: everything in row 0).
        Accumulates into `out` when given (that is what the multi-GPU all-reduce sums).  `index_out` (int32 [N]) receives the
        zone index of every point (steady-state loops: no allocation per call); `return_index` allocates it."""
        torch = _lib.require_cuda()
        Z = len(self.zones)
        if out is None:
            out = torch.zeros((num_slots, Z + 1), dtype=torch.int32, device=points.device)
        elif tuple(out.shape) != (num_slots, Z + 1) or out.dtype != torch.int32 or not out.is_contiguous():
            raise ValueError("out must be a contiguous int32 [num_slots, Z+1] tensor")
        idx, _ = self._launch(points, want_idx=return_index, hist=out, slot=slot, T=num_slots, transformer=transformer,
                              idx_out=index_out)
        return (out, idx) if (return_index or index_out is not None) else out

    def counts_to_dicts(self, hist) -> list[dict[str, int]]:
        """Dense [T, Z+1] histogram -> the reference's sparse per-frame dicts (non-zero bins only)."""
        h = hist.cpu().numpy()
        names = [*self._ids, "unclassified"]
        return [{names[j]: int(c) for j, c in enumerate(row) if c} for row in h]

    # -- reference surface ------------------------------------------------------------------------------
    def _ids_from_mask(self, m: int) -> list[str]:
        return [zid for z, zid in enumerate(self._ids) if (m >> z) & 1]

    def classify_batch(self, floor_points: Sequence[tuple[float, float]]) -> list[list[str]]:
        """Zone ids of every point, one launch (zone_classifier.py:151-160)."""
        if len(floor_points) == 0:
            return []
        torch = _lib.require_cuda()
        arr = np.array([(float(p[0]), float(p[1])) for p in floor_points], dtype=np.float64).reshape(-1, 2)
        pts = torch.from_numpy(arr).to(torch.device("cuda", torch.cuda.current_device()))
        _, mask = self._launch(pts, want_mask=True)
        words = mask.cpu().numpy().astype(np.uint64).reshape(len(arr), -1)      # [N, G] (G = 1 up to 64 zones)
        M = _lib.OPD_MAX_ZONES
        if self.allow_overlap:
            return [self._ids_from_mask(sum(int(w) << (M * g) for g, w in enumerate(row))) for row in words]
        # single label: every group's table already picked its own winner; keep the best of them by the global rank
        out = []
        for row in words:
            hits = [M * g + int(w).bit_length() - 1 for g, w in enumerate(row) if w]
            out.append([self._ids[min(hits, key=lambda z: self._rank[z])]] if hits else [])
        return out

    def classify(self, floor_point: tuple[float, float]) -> list[str]:
        """Zone ids containing the point; [] when outside every zone (zone_classifier.py:114-149)."""
        return self.classify_batch([floor_point])[0]

    def classify_with_unclassified(self, floor_point: tuple[float, float]) -> list[str]:
        return self.classify(floor_point) or ["unclassified"]

    def get_zone_info(self, zone_id: str) -> dict | None:
        for zone in self.zones:
            if zone["id"] == zone_id:
                return zone
        return None

    def get_all_zone_ids(self) -> list[str]:
        return list(self._ids)

    def get_zone_count(self) -> int:
        return len(self.zones)
