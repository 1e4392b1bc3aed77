"""CPU: pins oracle/pwa_oracle.py against the reference's OWN PiecewiseAffineTransformer outputs
(tests/golden/pwa_golden.npz, made by tests/golden/make_pwa_golden.py from src/transform/piecewise_affine.py)."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import pwa_oracle as po

from .conftest import GOLDEN

SCALE = (28.1926406926406, 28.241430700447)
MAP = (1878, 1369)


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN / "pwa_golden.npz"))


def test_oracle_matches_reference(golden):
    tb = po.build(golden["src"], golden["dst"])
    assert len(tb["simplices"]) == int(golden["num_triangles"])
    px, mm, within, tri, ext = po.transform_points(tb, golden["points"], scale_mm=SCALE, map_size=MAP)
    assert (ext == golden["extrapolated"]).all() and int(ext.sum()) > 50
    # a point ON a shared edge or vertex (the 24 correspondences themselves) belongs to several triangles whose lstsq affine maps
    # agree there only to ~1e-9 px (the reference's own training RMSE): 1e-9 away from edges, 1e-7 on them
    far = po.edge_distance(tb, golden["points"]) > 1e-6
    assert far.sum() > 450 and (tri[far] == golden["tri"][far]).all()
    np.testing.assert_allclose(px[far], golden["px"][far], rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(px, golden["px"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(mm, golden["mm"], rtol=0, atol=1e-5)
    assert (within == golden["within"]).all()
    bpx, _, _, btri, _ = po.transform_points(tb, golden["boxes"], is_bbox=True)
    np.testing.assert_allclose(bpx, golden["box_px"], rtol=1e-12, atol=1e-9)
    assert (btri == golden["box_tri"]).mean() > 0.98


def test_training_points_map_onto_their_targets(golden):
    tb = po.build(golden["src"], golden["dst"])
    px, *_ = po.transform_points(tb, golden["src"])
    np.testing.assert_allclose(px, golden["dst"], atol=1e-7)   # piecewise_affine.py docstring: RMSE 0 on the training data


def test_tps_oracle_matches_reference(golden):
    """oracle.tps_* == the reference's ThinPlateSplineTransformer (bit for bit: same solves, same summation order)."""
    tb = po.tps_build(golden["src"], golden["dst"])
    np.testing.assert_array_equal(po.tps_transform_points(tb, golden["points"][:200]), golden["tps_px"])
    np.testing.assert_array_equal(po.tps_transform_points(tb, golden["boxes"][:40], is_bbox=True), golden["tps_box_px"])
    np.testing.assert_allclose(po.tps_transform_points(tb, golden["src"]), golden["dst"], atol=1e-7)   # exact interpolation
