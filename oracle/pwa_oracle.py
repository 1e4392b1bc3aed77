"""TEST INFRASTRUCTURE ONLY (never imported by the product path): NumPy restatement of the reference's piecewise-affine
transform, pinned against the reference's own PiecewiseAffineTransformer in tests/golden/pwa_golden.npz
(tests/golden/make_pwa_golden.py imports src/transform/piecewise_affine.py).

  build(src, dst)          src/transform/piecewise_affine.py:59-125   Delaunay triangulation + one lstsq affine per triangle
  find_triangle            :127-136 -> scipy.spatial.Delaunay.find_simplex; restated as the brute-force form of scipy's
                           qhull.pyx _barycentric_inside on Delaunay.transform, eps = 100 * DBL_EPSILON, first hit in index order
  nearest_triangle         :138-153  argmin of the distance to the triangle centroids
  transform_points         :155-205  A @ [x, y, 1], bounds, mm scale; :207-236 foot point of a box
"""

from __future__ import annotations

import numpy as np

EPS = 100.0 * np.finfo(np.float64).eps


def build(src_points, dst_points) -> dict:
    from scipy.spatial import Delaunay

    src = np.array(src_points, dtype=np.float64)
    dst = np.array(dst_points, dtype=np.float64)
    tri = Delaunay(src)
    mats = []
    for simplex in tri.simplices:
        src_aug = np.vstack([src[simplex].T, np.ones(3)])
        dst_aug = np.vstack([dst[simplex].T, np.ones(3)])
        A, _, _, _ = np.linalg.lstsq(src_aug.T, dst_aug.T, rcond=None)
        mats.append(A.T)
    return {"src": src, "dst": dst, "simplices": tri.simplices.copy(), "transform": tri.transform.copy(),
            "affine": np.array(mats), "centroids": np.mean(src[tri.simplices], axis=1)}


def find_triangle(tb: dict, x: float, y: float) -> int:
    for t, tr in enumerate(tb["transform"]):
        dx, dy = x - tr[2, 0], y - tr[2, 1]
        c0 = tr[0, 0] * dx + tr[0, 1] * dy
        c1 = tr[1, 0] * dx + tr[1, 1] * dy
        c2 = 1.0 - c0 - c1
        if all(-EPS <= c <= 1.0 + EPS for c in (c0, c1, c2)):
            return t
    return -1


def nearest_triangle(tb: dict, x: float, y: float) -> int:
    return int(np.argmin(np.linalg.norm(tb["centroids"] - np.array([x, y]), axis=1)))


def transform_points(tb: dict, pts, is_bbox: bool = False, scale_mm=(1.0, 1.0), map_size=(np.inf, np.inf)):
    """-> (floor_px [N,2], floor_mm [N,2], within [N] bool, tri [N] int, extrapolated [N] bool)."""
    pts = np.asarray(pts, dtype=np.float64)
    n = len(pts)
    px, tri, ext = np.zeros((n, 2)), np.zeros(n, dtype=np.int64), np.zeros(n, dtype=bool)
    for i, row in enumerate(pts):
        x, y = (row[0] + row[2] / 2, row[1] + row[3]) if is_bbox else (row[0], row[1])
        t = find_triangle(tb, x, y)
        ext[i] = t < 0
        if t < 0:
            t = nearest_triangle(tb, x, y)
        tri[i] = t
        out = tb["affine"][t] @ np.array([x, y, 1.0])
        px[i] = out[:2]
    mm = px * np.array(scale_mm)
    within = (0 <= px[:, 0]) & (px[:, 0] < map_size[0]) & (0 <= px[:, 1]) & (px[:, 1] < map_size[1])
    return px, mm, within, tri, ext


def edge_distance(tb: dict, pts) -> np.ndarray:
    """Distance of every point to the nearest triangle edge (the triangle index is only pinned away from edges)."""
    pts = np.asarray(pts, dtype=np.float64)
    src, d = tb["src"], np.full(len(pts), np.inf)
    for s in tb["simplices"]:
        for a, b in ((0, 1), (1, 2), (2, 0)):
            p, q = src[s[a]], src[s[b]]
            v, w = q - p, pts - p
            t = np.clip((w @ v) / (v @ v), 0.0, 1.0)
            d = np.minimum(d, np.linalg.norm(w - t[:, None] * v, axis=1))
    return d


# ---- thin-plate spline (src/transform/piecewise_affine.py:398-545) -------------------------------------------------------
def tps_build(src_points, dst_points, regularization: float = 0.0) -> dict:
    """:445-485: kernel matrix of U(r) = r^2 log r, [K + lambda I, P; P^T, 0] solved once per output coordinate."""
    src = np.array(src_points, dtype=np.float64)
    dst = np.array(dst_points, dtype=np.float64)
    n = len(src)
    K = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                r = np.linalg.norm(src[i] - src[j])
                K[i, j] = r ** 2 * np.log(r) if r > 0 else 0.0
    P = np.hstack([np.ones((n, 1)), src])
    L = np.zeros((n + 3, n + 3))
    L[:n, :n] = K + regularization * np.eye(n)
    L[:n, n:] = P
    L[n:, :n] = P.T
    vx, vy = np.zeros(n + 3), np.zeros(n + 3)
    vx[:n], vy[:n] = dst[:, 0], dst[:, 1]
    cx, cy = np.linalg.solve(L, vx), np.linalg.solve(L, vy)
    return {"src": src, "wx": cx[:n], "wy": cy[:n], "ax": cx[n:], "ay": cy[n:]}


def tps_transform_points(tb: dict, pts, is_bbox: bool = False) -> np.ndarray:
    """:487-527: the radial-basis sum in index order with Python floats, then the affine part left to right."""
    pts = np.asarray(pts, dtype=np.float64)
    out = np.zeros((len(pts), 2))
    for k, row in enumerate(pts):
        x, y = (row[0] + row[2] / 2, row[1] + row[3]) if is_bbox else (row[0], row[1])
        rx = ry = 0.0
        for i, c in enumerate(tb["src"]):
            r = float(np.linalg.norm(np.array([x, y]) - c))
            u = r ** 2 * float(np.log(r)) if r > 0 else 0.0
            rx += float(tb["wx"][i]) * u
            ry += float(tb["wy"][i]) * u
        out[k] = (float(tb["ax"][0]) + float(tb["ax"][1]) * x + float(tb["ax"][2]) * y + rx,
                  float(tb["ay"][0]) + float(tb["ay"][1]) * x + float(tb["ay"][2]) * y + ry)
    return out
