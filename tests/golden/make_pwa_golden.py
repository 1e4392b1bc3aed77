"""Generates tests/golden/pwa_golden.npz with the reference's OWN PiecewiseAffineTransformer
(src/transform/piecewise_affine.py) on seeded correspondences: interior points, points outside the triangulation
(nearest-centroid extrapolation), the correspondence points themselves (training error 0) and boxes through
transform_batch.  Run here (CPU, needs /root/reference):  python tests/golden/make_pwa_golden.py"""

import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, "/root/reference")
from src.transform.floormap_config import FloorMapConfig  # noqa: E402
from src.transform.piecewise_affine import PiecewiseAffineTransformer, ThinPlateSplineTransformer  # noqa: E402


def main():
    rng = np.random.default_rng(17)
    # 24 correspondences: a jittered grid over the camera frame -> a smooth non-affine warp onto the floormap
    gx, gy = np.meshgrid(np.linspace(80, 1200, 6), np.linspace(120, 680, 4))
    src = np.stack([gx.ravel(), gy.ravel()], axis=1) + rng.uniform(-25, 25, (24, 2))
    dst = np.stack([200 + 1.2 * src[:, 0] + 0.0004 * src[:, 1] ** 2, 100 + 1.5 * src[:, 1] + 0.0003 * src[:, 0] ** 2], axis=1)
    fm = FloorMapConfig(width_px=1878, height_px=1369, origin_x_px=7, origin_y_px=9,
                        scale_x_mm_per_px=28.1926406926406, scale_y_mm_per_px=28.241430700447)
    tr = PiecewiseAffineTransformer(src, dst, fm)
    pts = np.concatenate([rng.uniform([0, 0], [1280, 720], (400, 2)), rng.uniform([-300, -200], [1600, 900], (100, 2)), src])
    res = [tr.transform_pixel((float(p[0]), float(p[1]))) for p in pts]
    boxes = np.concatenate([rng.uniform([0, 0], [1200, 600], (60, 2)), rng.uniform([10, 20], [200, 300], (60, 2))], axis=1)
    bres = tr.transform_batch([tuple(float(v) for v in b) for b in boxes])
    info = tr.get_info()
    tps = ThinPlateSplineTransformer(src, dst, fm)
    tres = [tps.transform_pixel((float(p[0]), float(p[1]))) for p in pts[:200]]
    tbres = tps.transform_batch([tuple(float(v) for v in b) for b in boxes[:40]])
    np.savez_compressed(
        Path(__file__).resolve().parent / "pwa_golden.npz", src=src, dst=dst, points=pts,
        px=np.array([r.floor_coords_px for r in res]), mm=np.array([r.floor_coords_mm for r in res]),
        within=np.array([r.is_within_bounds for r in res]), tri=np.array([r.triangle_index for r in res]),
        extrapolated=np.array([r.is_extrapolated for r in res]), boxes=boxes,
        box_px=np.array([r.floor_coords_px for r in bres]), box_tri=np.array([r.triangle_index for r in bres]),
        tps_px=np.array([r.floor_coords_px for r in tres]), tps_within=np.array([r.is_within_bounds for r in tres]),
        tps_box_px=np.array([r.floor_coords_px for r in tbres]), tps_rmse=np.array(tps.get_info()["training_error"]["rmse"]),
        num_triangles=np.array(info["num_triangles"]), rmse=np.array(info["training_error"]["rmse"]))
    print(len(pts), "points,", int(np.sum([r.is_extrapolated for r in res])), "extrapolated,", info["num_triangles"], "triangles, rmse",
          info["training_error"]["rmse"])


if __name__ == "__main__":
    main()
