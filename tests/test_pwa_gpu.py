"""GPU: PiecewiseAffineTransformer (csrc/pwa.cu through the C ABI) against the reference's own outputs
(tests/golden/pwa_golden.npz) and the NumPy oracle: floor coordinates to 1e-9 px (float64 fma vs NumPy's dot), the
extrapolation flag and bounds exactly, the triangle index exactly away from triangle edges; the reference's object surface
(transform_pixel / transform_detection / transform_batch / get_info / save / load) and the tensor flow into the zone kernel."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import floor_oracle as fo
from oracle import pwa_oracle as po

from .conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(GOLDEN / "pwa_golden.npz"))


@pytest.fixture(scope="module")
def transformer(golden, built_lib):
    import torch

    from office_person_detection_vit_b200.transform import FloorMapConfig, PiecewiseAffineTransformer

    torch.cuda.init()
    return PiecewiseAffineTransformer(golden["src"], golden["dst"], FloorMapConfig())


def test_points_match_reference_and_oracle(golden, transformer):
    import torch

    pts = torch.from_numpy(golden["points"]).cuda()
    px, mm, inb, tri, ext = (t.cpu().numpy() for t in transformer.transform_points(pts, with_mm=True, with_bounds=True, with_triangles=True))
    np.testing.assert_allclose(px, golden["px"], rtol=0, atol=1e-7)             # on shared edges / vertices the maps agree to ~1e-9
    np.testing.assert_allclose(mm, golden["mm"], rtol=0, atol=1e-5)
    assert (ext.astype(bool) == golden["extrapolated"]).all()
    assert (inb.astype(bool) == golden["within"]).all()
    tb = po.build(golden["src"], golden["dst"])
    o_px, _, _, o_tri, o_ext = po.transform_points(tb, golden["points"])
    assert (tri == o_tri).all() and (ext.astype(bool) == o_ext).all()          # same restatement: identical decisions
    far = po.edge_distance(tb, golden["points"]) > 1e-6
    assert (tri[far] == golden["tri"][far]).all()                               # the reference's find_simplex away from edges
    np.testing.assert_allclose(px[far], golden["px"][far], rtol=1e-12, atol=1e-9)


def test_reference_surface(golden, transformer, tmp_path):
    from office_person_detection_vit_b200.transform import FloorMapConfig, PiecewiseAffineTransformer

    res = transformer.transform_batch([tuple(float(v) for v in b) for b in golden["boxes"]])
    assert len(res) == len(golden["boxes"]) and all(r.is_valid for r in res)
    np.testing.assert_allclose(np.array([r.floor_coords_px for r in res]), golden["box_px"], rtol=1e-12, atol=1e-9)
    one = transformer.transform_detection(tuple(float(v) for v in golden["boxes"][3]))
    assert one.floor_coords_px == pytest.approx(tuple(golden["box_px"][3]), abs=1e-9) and one.floor_coords_mm is not None
    p = transformer.transform_pixel((float(golden["points"][0, 0]), float(golden["points"][0, 1])))
    assert p.floor_coords_px == pytest.approx(tuple(golden["px"][0]), abs=1e-9) and p.triangle_index == int(golden["tri"][0])
    assert transformer.transform_batch([]) == []
    info = transformer.get_info()
    assert info["method"] == "piecewise_affine" and info["num_triangles"] == int(golden["num_triangles"]) and info["num_points"] == 24
    assert info["training_error"]["rmse"] < 1e-6 and info["training_error"]["num_points"] == 24
    transformer.save(tmp_path / "pwa.pkl")
    again = PiecewiseAffineTransformer.load(tmp_path / "pwa.pkl", FloorMapConfig())
    assert again.transform_pixel((640.0, 360.0)).floor_coords_px == transformer.transform_pixel((640.0, 360.0)).floor_coords_px
    with pytest.raises(ValueError, match="最低3点"):
        PiecewiseAffineTransformer(golden["src"][:2], golden["dst"][:2])
    with pytest.raises(ValueError, match="一致しません"):
        PiecewiseAffineTransformer(golden["src"], golden["dst"][:5])


def test_pwa_points_feed_the_zone_kernel(golden, transformer):
    """PWA variant of Phase 3 on tensors: transform_points -> ZoneClassifier.count (no projection) == oracle counts."""
    import torch

    from office_person_detection_vit_b200.zone import ZoneClassifier

    zones = fo.grid_zones(16)
    zc = ZoneClassifier(zones, allow_overlap=False)
    rng = np.random.default_rng(5)
    pts = rng.uniform([0, 0], [1280, 720], (20000, 2))
    px = transformer.transform_points(torch.from_numpy(pts).cuda())
    hist, idx = zc.count(px, return_index=True)
    tb = po.build(golden["src"], golden["dst"])
    o_px, *_ = po.transform_points(tb, pts[:2000])
    np.testing.assert_allclose(px[:2000].cpu().numpy(), o_px, rtol=1e-12, atol=1e-9)
    o_idx, _ = fo.classify(px.cpu().numpy(), zones)
    assert (idx.cpu().numpy() == o_idx).all()
    assert (hist.cpu().numpy()[0] == fo.count(zone_idx=o_idx, Z=16)[0]).all()


def test_lookup_grid_equals_full_search(golden, transformer):
    """The cell lookup (strictly-inside cells, outside cells with a unique nearest centroid) must reproduce the full search
    bit for bit: floor coordinates, triangle index and extrapolation flag, inside, around and far outside the grid."""
    import torch

    from office_person_detection_vit_b200 import _lib

    rng = np.random.default_rng(23)
    pts = np.concatenate([rng.uniform([-700, -500], [2000, 1300], (150000, 2)), rng.uniform([-1e5, -1e5], [1e5, 1e5], (2000, 2)),
                          golden["src"], np.array([[np.nan, 1.0], [np.inf, -np.inf], [0.0, 0.0]])])
    t = torch.from_numpy(pts).cuda()
    a = transformer.transform_points(t, with_triangles=True)
    try:
        _lib.check(_lib.lib().opd_set_option(b"probe", 64), "opd_set_option")
        b = transformer.transform_points(t, with_triangles=True)
    finally:
        _lib.lib().opd_set_option(b"probe", 0)
    torch.cuda.synchronize()
    for x, y in zip(a, b):
        assert torch.equal(x.view(torch.int64) if x.dtype == torch.float64 else x, y.view(torch.int64) if y.dtype == torch.float64 else y)


def test_thin_plate_spline(golden, built_lib):
    """ThinPlateSplineTransformer (tps_transform_kernel) against the reference's own outputs: same coefficients (host solve as
    in the reference), same summation order, float64; CUDA's log() may differ from NumPy's by an ulp per term, which the
    24-term sum of magnitudes ~1e6 turns into <= 1e-8 px."""
    import torch

    from office_person_detection_vit_b200.transform import FloorMapConfig, ThinPlateSplineTransformer

    tps = ThinPlateSplineTransformer(golden["src"], golden["dst"], FloorMapConfig())
    px, mm, inb = (t.cpu().numpy() for t in tps.transform_points(torch.from_numpy(golden["points"][:200]).cuda(), with_mm=True, with_bounds=True))
    np.testing.assert_allclose(px, golden["tps_px"], rtol=0, atol=1e-8)
    assert (inb.astype(bool) == golden["tps_within"]).all()
    np.testing.assert_allclose(mm, px * np.array([28.1926406926406, 28.241430700447]), rtol=1e-15)
    res = tps.transform_batch([tuple(float(v) for v in b) for b in golden["boxes"][:40]])
    np.testing.assert_allclose(np.array([r.floor_coords_px for r in res]), golden["tps_box_px"], rtol=0, atol=1e-8)
    info = tps.get_info()
    assert info["method"] == "thin_plate_spline" and info["num_points"] == 24 and info["training_error"]["rmse"] < 1e-6
    assert tps.transform_batch([]) == []
    with pytest.raises(ValueError, match="最低3点"):
        ThinPlateSplineTransformer(golden["src"][:2], golden["dst"][:2])
