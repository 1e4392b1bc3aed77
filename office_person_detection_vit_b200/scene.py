"""Real-scene constants of the reference and the synthetic workload generators of SURVEY.md §8d.

Data only (no projection / classification arithmetic lives here): shared by bench.py, benchmarks/ and — through
oracle/floor_oracle.py — by both sides of every parity test."""

from __future__ import annotations

import math

import numpy as np

# config.yaml:115-119 (homography.matrix) and :208-223 (floormap) of the reference
H_CONFIG = np.array([
    [-0.8795888447, -2.8974379541, 417.8510123786],
    [-1.5459702925, -3.4570021203, 1054.0107447082],
    [-0.0011928509, -0.0035480452, 1.0000000000],
], dtype=np.float64)
MAP_W, MAP_H = 1878, 1369
SX_MM, SY_MM = 28.1926406926406, 28.241430700447


def grid_zones(Z: int, width=MAP_W, height=MAP_H) -> list[dict]:
    """Z axis-aligned cells of a g x g grid over the floormap, g = ceil(sqrt(Z)); priority = index."""
    g = math.ceil(math.sqrt(Z))
    cw, ch = width / g, height / g
    zones = []
    for i in range(Z):
        cx, cy = i % g, i // g
        x0, y0, x1, y1 = cx * cw, cy * ch, (cx + 1) * cw, (cy + 1) * ch
        zones.append({"id": f"zone_{i + 1}", "polygon": [[x0, y0], [x1, y0], [x1, y1], [x0, y1]], "priority": i})
    return zones


def star_zones(Z: int, seed: int, width=MAP_W, height=MAP_H) -> list[dict]:
    """Per grid cell a random 5-8 vertex, possibly concave, rotated polygon inset in the cell, plus 3
    deliberately overlapping pairs with distinct priorities (every 5th zone has priority None)."""
    rng = np.random.default_rng(seed)
    g = math.ceil(math.sqrt(Z))
    cw, ch = width / g, height / g
    zones = []
    for i in range(Z):
        cx, cy = (i % g + 0.5) * cw, (i // g + 0.5) * ch
        nv = int(rng.integers(5, 9))
        ang = np.sort(rng.uniform(0, 2 * math.pi, nv)) + rng.uniform(0, 2 * math.pi)
        rad = rng.uniform(0.25, 0.48, nv)
        poly = [[float(cx + r * cw * math.cos(a)), float(cy + r * ch * math.sin(a))] for a, r in zip(ang, rad)]
        zones.append({"id": f"zone_{i + 1}", "polygon": poly, "priority": None if i % 5 == 4 else float(Z - i)})
    for k in range(min(3, Z // 2)):  # overlapping pairs: shift a copy of zone 2k over zone 2k+1
        a, b = zones[2 * k], zones[2 * k + 1]
        dx = (b["polygon"][0][0] - a["polygon"][0][0]) * 0.6
        b["polygon"] = [[x + dx * 0.1, y] for x, y in b["polygon"]]
        a["polygon"] = [[x + dx, y] for x, y in a["polygon"]]
    return zones


def camera_points(n: int, seed: int) -> np.ndarray:
    """Camera-space points x~U[0,1280), y~U[0,720), float32."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(0, 1280, n), rng.uniform(0, 720, n)], axis=1).astype(np.float32)
