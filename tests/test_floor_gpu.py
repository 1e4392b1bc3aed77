"""GPU parity of K9-K11 (csrc/floor.cu) through the C ABI and the reference-shaped Python classes,
against (a) outputs of the reference itself (tests/golden) and (b) the CPU oracle on seeded inputs."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import floor_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    torch.cuda.init()
    return torch


def _engines(H, zones, allow_overlap):
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer
    from office_person_detection_vit_b200.zone import ZoneClassifier

    return HomographyTransformer(H, FloorMapConfig()), ZoneClassifier(zones, allow_overlap=allow_overlap)


# ---- reference-shaped surface vs the reference's own outputs -----------------------------------------
def test_transform_batch_matches_reference(golden, torch_cuda, built_lib):
    _, g, zones = golden
    tr, _ = _engines(g["H"], zones, False)
    res = tr.transform_batch([tuple(map(float, b)) for b in g["boxes"]])
    px = np.array([r.floor_coords_px for r in res])
    mm = np.array([r.floor_coords_mm for r in res])
    np.testing.assert_allclose(px, g["floor_px"], rtol=1e-12, atol=1e-9)   # float64, <= a few ulp (FMA vs BLAS)
    np.testing.assert_allclose(mm, g["floor_mm"], rtol=1e-12, atol=1e-7)
    assert [r.is_within_bounds for r in res] == [bool(w) for w in g["within"]]
    assert all(r.is_valid for r in res)
    assert tr.transform_batch([]) == []


def test_classify_matches_reference(golden, torch_cuda, built_lib):
    _, g, zones = golden
    ids = [z["id"] for z in zones]
    _, zc1 = _engines(g["H"], zones, False)
    _, zcm = _engines(g["H"], zones, True)
    pts = [tuple(map(float, p)) for p in g["floor_px"]]
    single = zc1.classify_batch(pts)
    multi = zcm.classify_batch(pts)
    exp_single = [[ids[i]] if i >= 0 else [] for i in g["zone_idx"]]
    exp_multi = [[ids[z] for z in range(len(ids)) if (int(m) >> z) & 1] for m in g["zone_mask"]]
    assert single == exp_single      # bit exact on the reference's own float64 floor points
    assert multi == exp_multi
    assert zc1.classify(pts[0]) == exp_single[0]
    assert zc1.classify_with_unclassified((1e7, 1e7)) == ["unclassified"]


def test_fused_count_matches_reference(golden, torch_cuda, built_lib):
    """boxes -> foot -> H -> zone -> per-frame histogram in one launch, float64 entry."""
    torch = torch_cuda
    _, g, zones = golden
    tr, zc1 = _engines(g["H"], zones, False)
    _, zcm = _engines(g["H"], zones, True)
    boxes = torch.from_numpy(g["boxes"]).cuda()
    foot = torch.stack([boxes[:, 0] + boxes[:, 2] / 2, boxes[:, 1] + boxes[:, 3]], dim=1).contiguous()
    slot = torch.from_numpy(g["frame_of"]).cuda()
    T = g["hist_single"].shape[0]
    h1, idx = zc1.count(foot, slot=slot, num_slots=T, transformer=tr, return_index=True)
    hm = zcm.count(foot, slot=slot, num_slots=T, transformer=tr)
    assert (idx.cpu().numpy() == g["zone_idx"]).all()
    assert (h1.cpu().numpy() == g["hist_single"]).all()
    assert (hm.cpu().numpy() == g["hist_multi"]).all()
    assert zc1.counts_to_dicts(h1)[0] == {
        ([z["id"] for z in zones] + ["unclassified"])[j]: int(c) for j, c in enumerate(g["hist_single"][0]) if c}


def test_aggregator_counts(torch_cuda, built_lib):
    """tests/test_aggregator.py:28-56 of the reference: multi-zone and unclassified detections."""
    from office_person_detection_vit_b200.aggregation import Aggregator
    from office_person_detection_vit_b200.models import Detection

    def det(zs):
        return Detection(bbox=(0, 0, 1, 1), confidence=0.9, class_id=1, class_name="person", camera_coords=(0, 0),
                         zone_ids=zs)

    agg = Aggregator()
    counts = agg.aggregate_frame("12:00", [det(["zone_a"]), det(["zone_a", "zone_b"]), det([]), det(["zone_b"])])
    assert counts == {"zone_a": 2, "zone_b": 2, "unclassified": 1}
    assert list(counts) == ["zone_a", "zone_b", "unclassified"]
    assert agg.get_zone_counts([]) == {}
    assert len(agg.results) == 3


# ---- tensor entries vs the oracle on seeded inputs ------------------------------------------------------
@pytest.mark.parametrize("Z,kind", [(4, "grid"), (16, "star"), (64, "star"), (64, "grid")])
def test_fast_kernel_matches_oracle(Z, kind, torch_cuda, built_lib):
    """float32 camera points, 2^20 points: the filtered kernel (float32 filter + float64 queue) must give
    the oracle's zone index everywhere outside the 1e-4 px edge band, and the exact histogram of its own
    indices."""
    torch = torch_cuda
    zones = fo.grid_zones(Z) if kind == "grid" else fo.star_zones(Z, seed=Z // 4)
    n = 1 << 20
    pts = fo.camera_points(n, seed=3)
    tr, zc = _engines(fo.H_CONFIG, zones, False)
    d_pts = torch.from_numpy(pts).cuda()
    hist, idx = zc.count(d_pts, transformer=tr, return_index=True)
    idx = idx.cpu().numpy()
    exp_idx, exp_hist = fo.project_classify_count(fo.H_CONFIG, pts, zones)
    px, _, _ = fo.transform(fo.H_CONFIG, pts.astype(np.float64), is_bbox=False)
    far = fo.min_edge_distance(px, zones) >= 1e-4
    assert (idx[far] == exp_idx[far]).all()
    assert (idx != exp_idx).sum() <= 2          # float64 path: in practice identical everywhere
    mine = fo.count(zone_idx=idx, Z=Z)[0]
    assert (hist.cpu().numpy()[0] == mine).all()
    if (idx == exp_idx).all():
        assert (hist.cpu().numpy()[0] == exp_hist).all()
    # count-only launch (no index written) gives the same histogram
    hist2 = zc.count(d_pts, transformer=tr)
    assert torch.equal(hist, hist2)
    # exact kernel (forced by asking for masks) agrees with the fast kernel
    masks = zc.classify_masks(d_pts, transformer=tr).cpu().numpy().astype(np.uint64)
    win = np.full(n, -1, dtype=np.int32)
    for z in range(Z):
        win[masks == np.uint64(1 << z)] = z
    assert ((masks & (masks - np.uint64(1))) == 0).all()   # single-label table: at most one bit
    assert (win == idx).all()


def test_fast_kernel_ragged_tail_and_idempotence(torch_cuda, built_lib):
    torch = torch_cuda
    zones = fo.grid_zones(16)
    tr, zc = _engines(fo.H_CONFIG, zones, False)
    for n in (65536, 65537, 100003, 4096 * 37 + 5):
        pts = fo.camera_points(n, seed=n)
        d = torch.from_numpy(pts).cuda()
        hist, idx = zc.count(d, transformer=tr, return_index=True)
        exp_idx, exp_hist = fo.project_classify_count(fo.H_CONFIG, pts, zones)
        assert (idx.cpu().numpy() == exp_idx).all()
        assert (hist.cpu().numpy()[0] == exp_hist).all()
        assert int(hist.sum()) == n
        again = zc.count(d, transformer=tr, out=hist.clone())   # accumulates: exactly doubles
        assert torch.equal(again, 2 * hist)


def test_fast_kernel_overlapping_zones(torch_cuda, built_lib):
    """allow_overlap=True (the class default) on zones that really overlap: the filtered kernel (N >= 65536, no slots) must count a
    point once in EVERY containing zone (aggregator.py:66-69), like the exact kernel and the oracle - ADVICE r1."""
    torch = torch_cuda
    zones = fo.star_zones(16, seed=2)               # 3 deliberately overlapping pairs
    zones += [{"id": "big", "polygon": [[300.0, 200.0], [1500.0, 250.0], [1400.0, 1100.0], [350.0, 1000.0]], "priority": 99}]
    Z = len(zones)
    for allow in (True, False):
        tr, zc = _engines(fo.H_CONFIG, zones, allow)
        for n in (1 << 17, (1 << 17) + 77):
            pts = fo.camera_points(n, seed=21)
            d = torch.from_numpy(pts).cuda()
            hist, idx = zc.count(d, transformer=tr, return_index=True)                       # fast kernel
            masks = zc.classify_masks(d, transformer=tr).cpu().numpy().astype(np.uint64)     # exact kernel
            slot = torch.zeros(n, dtype=torch.int32, device="cuda")
            hist_exact, idx_exact = zc.count(d, slot=slot, num_slots=1, transformer=tr, return_index=True)   # exact kernel
            assert torch.equal(idx, idx_exact)
            assert torch.equal(hist, hist_exact)
            exp = np.array([int(((masks >> np.uint64(z)) & np.uint64(1)).sum()) for z in range(Z)] + [int((masks == 0).sum())])
            assert (hist.cpu().numpy()[0] == exp).all()
            if allow:
                assert ((masks & (masks - np.uint64(1))) != 0).sum() > 1000      # the case under test really occurs
                assert int(hist.sum()) > n
            exp_idx, _ = fo.project_classify_count(fo.H_CONFIG, pts, zones)
            assert (idx.cpu().numpy() != exp_idx).sum() <= 2
            assert torch.equal(zc.count(d, transformer=tr), hist)                          # count-only launch


def test_unused_rows_are_not_classified(torch_cuda, built_lib):
    """Rows whose slot is outside [0, T) (the padding rows of the detector's [B, 100] tables: foot (0, 0), slot -1) get zone
    index -1 / mask 0 and no count, wherever H maps them."""
    torch = torch_cuda
    zones = [{"id": "all", "polygon": [[-1e6, -1e6], [1e6, -1e6], [1e6, 1e6], [-1e6, 1e6]]}]
    tr, zc = _engines(np.eye(3), zones, False)
    pts = torch.zeros(8, 2, dtype=torch.float64, device="cuda")
    slot = torch.tensor([0, -1, 1, -1, 5, 1, -1, 0], dtype=torch.int32, device="cuda")
    hist, idx = zc.count(pts, slot=slot, num_slots=2, transformer=tr, return_index=True)
    assert idx.cpu().tolist() == [0, -1, 0, -1, -1, 0, -1, 0]
    assert hist.cpu().tolist() == [[2, 0], [2, 0]]


def test_degenerate_points(torch_cuda, built_lib):
    """Points on the horizon (W = 0) and NaN/inf inputs are unclassified, as in the reference (every comparison
    with NaN is False); a huge finite point still projects to a finite ratio X/W and is classified like the oracle."""
    torch = torch_cuda
    zones = fo.grid_zones(4)
    tr, zc = _engines(fo.H_CONFIG, zones, False)
    n = 70000
    pts = fo.camera_points(n, seed=11)
    pts[0] = (np.nan, 5.0); pts[1] = (np.inf, 1.0); pts[2] = (1e30, -1e30)
    pts[3] = (0.0, 1.0 / 0.0035480452)          # W ~ 0
    d = torch.from_numpy(pts).cuda()
    idx = zc.classify_points(d, transformer=tr).cpu().numpy()
    with np.errstate(all="ignore"):
        exp_idx, _ = fo.project_classify_count(fo.H_CONFIG, pts, zones)
    assert (idx[:2] == -1).all()
    assert (idx == exp_idx).all()


def test_transform_points_f32_rounding(torch_cuda, built_lib):
    """float32 tensor entry = float64 arithmetic rounded once to float32."""
    torch = torch_cuda
    from office_person_detection_vit_b200.transform import FloorMapConfig, HomographyTransformer

    tr = HomographyTransformer(fo.H_CONFIG, FloorMapConfig())
    pts = fo.camera_points(5000, seed=5)
    px, mm, inb = tr.transform_points(torch.from_numpy(pts).cuda(), with_mm=True, with_bounds=True)
    epx, emm, ein = fo.transform(fo.H_CONFIG, pts.astype(np.float64), is_bbox=False)
    with np.errstate(all="ignore"):
        ok = np.isfinite(epx).all(axis=1) & (np.abs(epx) < 1e30).all(axis=1)
        np.testing.assert_allclose(px.cpu().numpy()[ok], epx[ok].astype(np.float32), rtol=2e-7, atol=0)
        np.testing.assert_allclose(mm.cpu().numpy()[ok], emm[ok].astype(np.float32), rtol=2e-7, atol=0)
    assert (inb.cpu().numpy() == ein).all()


def test_zone_table_info_and_limits(torch_cuda, built_lib):
    from office_person_detection_vit_b200.zone import ZoneClassifier

    zc = ZoneClassifier(fo.grid_zones(64), allow_overlap=False)
    info = zc.table.info()
    assert info["grid_w"] * info["grid_h"] <= 160 * 1024 and info["boundary_cells"] > 0
    assert info["boundary_cells"] < 0.12 * info["grid_w"] * info["grid_h"]


@pytest.mark.parametrize("allow_overlap", [False, True])
def test_more_than_64_zones(allow_overlap, torch_cuda, built_lib):
    """The reference sets no limit on the number of zones (zone_classifier.py:44-112): 100 zones = two device tables + the combine
    kernel.  Checked against the line-by-line Python restatement of classify (the C oracle's masks are 64-bit): zone ids per point
    through the object surface, the tensor index, and the [T, Z + 1] histogram with per-timestamp slots."""
    import math

    torch = torch_cuda
    from office_person_detection_vit_b200.zone import ZoneClassifier

    zones = fo.star_zones(100, seed=6)
    zones[70]["priority"] = 0.5          # a late zone that beats every earlier one where they overlap
    zones += [{"id": "wide", "polygon": [[200.0, 150.0], [1700.0, 180.0], [1650.0, 1200.0], [250.0, 1150.0]], "priority": 1e6}]
    Z = len(zones)
    zc = ZoneClassifier(zones, allow_overlap=allow_overlap)
    assert zc.grouped and zc.get_zone_count() == Z
    rng = np.random.default_rng(5)
    n = 3000
    pts = np.stack([rng.uniform(-50, 1950, n), rng.uniform(-50, 1420, n)], axis=1)
    slot = rng.integers(0, 3, n).astype(np.int32)
    ids = [z["id"] for z in zones]
    exp_ids, exp_idx = [], []
    for x, y in pts:
        hit = [i for i, z in enumerate(zones) if fo.point_in_polygon_py(float(x), float(y), z["polygon"])]
        win = min(hit, key=lambda i: (math.inf if zones[i]["priority"] is None else zones[i]["priority"], i)) if hit else -1
        exp_idx.append(win)
        exp_ids.append([ids[i] for i in hit] if allow_overlap else ([ids[win]] if hit else []))
    got = zc.classify_batch([tuple(p) for p in pts])
    assert got == exp_ids
    assert sum(len(e) > 1 for e in exp_ids) > 50 or not allow_overlap
    d = torch.from_numpy(pts).cuda()
    idx = zc.classify_points(d).cpu().numpy()
    assert (idx == np.array(exp_idx)).all()
    assert (idx >= 64).sum() > 100                                   # zones of the second table are really used
    hist = zc.count(d, slot=torch.from_numpy(slot).cuda(), num_slots=3).cpu().numpy()
    exp_hist = np.zeros((3, Z + 1), np.int64)
    for s_, e in zip(slot, exp_ids):
        for zid in e:
            exp_hist[s_, ids.index(zid)] += 1
        if not e:
            exp_hist[s_, Z] += 1
    assert (hist == exp_hist).all()
    assert zc.counts_to_dicts(torch.from_numpy(hist))[0] == {([*ids, "unclassified"])[j]: int(c) for j, c in enumerate(exp_hist[0]) if c}
